// ba_geometry.cuh -- BAL reprojection residual and analytical Jacobian, device side.
//
// Computes the same quantities as the reference's AnalyticalDerivativesKernelGradKernel
// (src/geo/analytical_derivatives.cu:16-285): r(2) and J(2x12) of one observation
// for a camera [omega(3), t(3), f, k1, k2] and a point X(3).  The derivation here is
// our own closed form rather than the reference's term-by-term product rule:
//
//   theta^2 = |omega|^2,  a = sin(theta)/theta,  b = (1-cos(theta))/theta^2
//   R  = c I + a [omega]x + b omega omega^T                 (c = cos(theta))
//   P  = R X + t
//   dP/dX     = R
//   dP/domega = V omega^T - a [X]x + b omega X^T + b (omega.X) I
//        with V = -a X + a' (omega x X) + b' (omega.X) omega,
//             a' = (c - a)/theta^2,  b' = (a - 2b)/theta^2
//   (theta^2 <= eps: R = I + [omega]x, dP/domega = -[X]x, as the reference :131-158)
//   p = -P_xy / P_z,  l2 = |p|^2,  d = 1 + k1 l2 + k2 l2^2,  r = f d p - obs
//   dr/dP is formed once per residual row as the 3-vector `ef`; then
//   J_row[0:3] = ef^T dP/domega, J_row[3:6] = ef, J_row[6:9] = (d p, f l2 p, f l2^2 p),
//   J_row[9:12] = ef^T R.
#pragma once
#include <cuda_runtime.h>

namespace ba {

template <typename T> struct Eps;
template <> struct Eps<double> { static __host__ __device__ constexpr double v() { return 2.220446049250313e-16; } };
template <> struct Eps<float> { static __host__ __device__ constexpr float v() { return 1.1920928955078125e-07f; } };

__host__ __device__ __forceinline__ void sincos_t(double x, double *s, double *c) { sincos(x, s, c); }
__host__ __device__ __forceinline__ void sincos_t(float x, float *s, float *c) { sincosf(x, s, c); }
__host__ __device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__host__ __device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }

// Per-camera rotation constants (depend on omega only).
template <typename T>
struct RotConst {
  T c, a, b, ap, bp;  // cos, sin/theta, (1-cos)/theta^2, a', b'
  bool small;
};

template <typename T>
__host__ __device__ __forceinline__ RotConst<T> rot_const(const T w0, const T w1, const T w2) {
  RotConst<T> k;
  const T th2 = w0 * w0 + w1 * w1 + w2 * w2;
  k.small = !(th2 > Eps<T>::v());
  if (!k.small) {
    const T th = sqrt_t(th2);
    T s;
    sincos_t(th, &s, &k.c);
    const T ith2 = T(1) / th2;
    k.a = s / th;
    k.b = (T(1) - k.c) * ith2;
    k.ap = (k.c - k.a) * ith2;
    k.bp = (k.a - T(2) * k.b) * ith2;
  } else {
    k.c = T(1); k.a = T(1); k.b = T(0); k.ap = T(0); k.bp = T(0);
  }
  return k;
}

// Residual only.  cam = pointer to 9 contiguous scalars (registers or memory).
template <typename T>
__host__ __device__ __forceinline__ void edge_residual(const T *cam, const RotConst<T> &k, const T X0, const T X1,
                                              const T X2, const T o0, const T o1, T &r0, T &r1) {
  const T w0 = cam[0], w1 = cam[1], w2 = cam[2];
  const T cx0 = w1 * X2 - w2 * X1, cx1 = w2 * X0 - w0 * X2, cx2 = w0 * X1 - w1 * X0;  // omega x X
  const T d = w0 * X0 + w1 * X1 + w2 * X2;
  T P0, P1, P2;
  if (!k.small) {
    const T bd = k.b * d;
    P0 = k.c * X0 + k.a * cx0 + bd * w0;
    P1 = k.c * X1 + k.a * cx1 + bd * w1;
    P2 = k.c * X2 + k.a * cx2 + bd * w2;
  } else {
    P0 = X0 + cx0; P1 = X1 + cx1; P2 = X2 + cx2;
  }
  P0 += cam[3]; P1 += cam[4]; P2 += cam[5];
  const T iz = T(1) / P2;
  const T B0 = -P0 * iz, B1 = -P1 * iz;
  const T l2 = B0 * B0 + B1 * B1;
  const T fr = cam[6] * (T(1) + cam[7] * l2 + cam[8] * l2 * l2);
  r0 = fr * B0 - o0;
  r1 = fr * B1 - o1;
}

// Residual + Jacobian.  J is [2][12] (row-major), columns 0-8 camera, 9-11 point.
template <typename T>
__host__ __device__ __forceinline__ void edge_residual_jacobian(const T *cam, const RotConst<T> &k, const T X0,
                                                       const T X1, const T X2, const T o0, const T o1,
                                                       T &r0, T &r1, T *J) {
  const T w0 = cam[0], w1 = cam[1], w2 = cam[2];
  const T cx0 = w1 * X2 - w2 * X1, cx1 = w2 * X0 - w0 * X2, cx2 = w0 * X1 - w1 * X0;
  const T d = w0 * X0 + w1 * X1 + w2 * X2;
  T P0, P1, P2, V0, V1, V2;
  if (!k.small) {
    const T bd = k.b * d;
    P0 = k.c * X0 + k.a * cx0 + bd * w0;
    P1 = k.c * X1 + k.a * cx1 + bd * w1;
    P2 = k.c * X2 + k.a * cx2 + bd * w2;
    const T bpd = k.bp * d;
    V0 = -k.a * X0 + k.ap * cx0 + bpd * w0;
    V1 = -k.a * X1 + k.ap * cx1 + bpd * w1;
    V2 = -k.a * X2 + k.ap * cx2 + bpd * w2;
  } else {
    P0 = X0 + cx0; P1 = X1 + cx1; P2 = X2 + cx2;
    V0 = V1 = V2 = T(0);
  }
  P0 += cam[3]; P1 += cam[4]; P2 += cam[5];
  const T f = cam[6], k1 = cam[7], k2 = cam[8];
  const T iz = T(1) / P2;
  const T B0 = -P0 * iz, B1 = -P1 * iz;
  const T l2 = B0 * B0 + B1 * B1;
  const T rp = T(1) + k1 * l2 + k2 * l2 * l2;
  const T fr = f * rp;
  r0 = fr * B0 - o0;
  r1 = fr * B1 - o1;
  const T g2 = T(2) * f * (k1 + T(2) * k2 * l2);  // 2 f d(rp)/d(l2)
  const T A00 = fr + g2 * B0 * B0, A01 = g2 * B0 * B1, A11 = fr + g2 * B1 * B1;
  const T bd = k.small ? T(0) : k.b * d;
#pragma unroll
  for (int row = 0; row < 2; ++row) {
    const T Ax = row == 0 ? A00 : A01, Ay = row == 0 ? A01 : A11;
    const T e0 = -iz * Ax, e1 = -iz * Ay, e2 = -iz * (Ax * B0 + Ay * B1);  // dr_row/dP
    T *Jr = J + 12 * row;
    // ef^T dP/domega = (ef.V) omega^T + a (X x ef)^T + b (ef.omega) X^T + b d ef^T
    const T eV = e0 * V0 + e1 * V1 + e2 * V2;
    const T ew = e0 * w0 + e1 * w1 + e2 * w2;
    const T xe0 = X1 * e2 - X2 * e1, xe1 = X2 * e0 - X0 * e2, xe2 = X0 * e1 - X1 * e0;  // X x ef
    const T bew = k.small ? T(0) : k.b * ew;
    Jr[0] = eV * w0 + k.a * xe0 + bew * X0 + bd * e0;
    Jr[1] = eV * w1 + k.a * xe1 + bew * X1 + bd * e1;
    Jr[2] = eV * w2 + k.a * xe2 + bew * X2 + bd * e2;
    Jr[3] = e0; Jr[4] = e1; Jr[5] = e2;
    const T Br = row == 0 ? B0 : B1;
    Jr[6] = rp * Br;
    Jr[7] = f * l2 * Br;
    Jr[8] = f * l2 * l2 * Br;
    // ef^T R = c ef^T + a (ef x omega)^T + b (ef.omega) omega^T
    const T ow0 = e1 * w2 - e2 * w1, ow1 = e2 * w0 - e0 * w2, ow2 = e0 * w1 - e1 * w0;  // ef x omega
    Jr[9] = k.c * e0 + k.a * ow0 + bew * w0;
    Jr[10] = k.c * e1 + k.a * ow1 + bew * w1;
    Jr[11] = k.c * e2 + k.a * ow2 + bew * w2;
  }
}

// Residual, point-side Jacobian Jp = dr/dX (2x3, row-major) and the normalised image point B = -P_xy/P_z.
// Same arithmetic as edge_residual_jacobian for these outputs (the factored operator stream stores exactly
// these eight numbers per observation, see ba_internal.h).
template <typename T>
__host__ __device__ __forceinline__ void edge_point_jacobian(const T *cam, const RotConst<T> &k, const T X0,
                                                    const T X1, const T X2, const T o0, const T o1,
                                                    T &r0, T &r1, T *Jp, T &B0o, T &B1o) {
  const T w0 = cam[0], w1 = cam[1], w2 = cam[2];
  const T cx0 = w1 * X2 - w2 * X1, cx1 = w2 * X0 - w0 * X2, cx2 = w0 * X1 - w1 * X0;
  const T d = w0 * X0 + w1 * X1 + w2 * X2;
  T P0, P1, P2;
  if (!k.small) {
    const T bd = k.b * d;
    P0 = k.c * X0 + k.a * cx0 + bd * w0;
    P1 = k.c * X1 + k.a * cx1 + bd * w1;
    P2 = k.c * X2 + k.a * cx2 + bd * w2;
  } else {
    P0 = X0 + cx0; P1 = X1 + cx1; P2 = X2 + cx2;
  }
  P0 += cam[3]; P1 += cam[4]; P2 += cam[5];
  const T f = cam[6], k1 = cam[7], k2 = cam[8];
  const T iz = T(1) / P2;
  const T B0 = -P0 * iz, B1 = -P1 * iz;
  const T l2 = B0 * B0 + B1 * B1;
  const T rp = T(1) + k1 * l2 + k2 * l2 * l2;
  const T fr = f * rp;
  r0 = fr * B0 - o0;
  r1 = fr * B1 - o1;
  B0o = B0; B1o = B1;
  const T g2 = T(2) * f * (k1 + T(2) * k2 * l2);
  const T A00 = fr + g2 * B0 * B0, A01 = g2 * B0 * B1, A11 = fr + g2 * B1 * B1;
#pragma unroll
  for (int row = 0; row < 2; ++row) {
    const T Ax = row == 0 ? A00 : A01, Ay = row == 0 ? A01 : A11;
    const T e0 = -iz * Ax, e1 = -iz * Ay, e2 = -iz * (Ax * B0 + Ay * B1);
    const T ew = e0 * w0 + e1 * w1 + e2 * w2;
    const T bew = k.small ? T(0) : k.b * ew;
    const T ow0 = e1 * w2 - e2 * w1, ow1 = e2 * w0 - e0 * w2, ow2 = e0 * w1 - e1 * w0;
    Jp[3 * row] = k.c * e0 + k.a * ow0 + bew * w0;
    Jp[3 * row + 1] = k.c * e1 + k.a * ow1 + bew * w1;
    Jp[3 * row + 2] = k.c * e2 + k.a * ow2 + bew * w2;
  }
}

// Per-camera blocks of the factored operator (ba_internal.h): out[0..8] = Jr (right Jacobian of SO(3) at
// omega: R^T dR(omega)[x] = [Jr x]x), out[9..17] = R, out[18..26] = E = [[1,0,0],[k1,f,0],[k2,0,f]], all row-major.
// Only meaningful for cameras outside the small-angle branch (the reference's first-order R = I + [omega]x is not
// a rotation; a solver with such a camera streams the full Jacobian instead).
template <typename T>
__host__ __device__ __forceinline__ void camera_transform(const T *cam, const RotConst<T> &k, T *out) {
  const T w0 = cam[0], w1 = cam[1], w2 = cam[2];
  const T th2 = w0 * w0 + w1 * w1 + w2 * w2;
  // c3 = (1 - a) / theta^2 = (theta - sin theta) / theta^3: series below theta^2 = 1e-4 (next term theta^6/362880)
  T c3, b = k.b;
  if (th2 < T(1e-4)) {
    c3 = T(1) / T(6) - th2 * (T(1) / T(120) - th2 * (T(1) / T(5040)));
    b = T(0.5) - th2 * (T(1) / T(24) - th2 * (T(1) / T(720)));
  } else {
    c3 = (T(1) - k.a) / th2;
  }
  // Jr = I - b [w]x + c3 [w]x^2,  [w]x^2 = w w^T - theta^2 I
  const T ww[9] = {w0 * w0, w0 * w1, w0 * w2, w1 * w0, w1 * w1, w1 * w2, w2 * w0, w2 * w1, w2 * w2};
  const T K[9] = {T(0), -w2, w1, w2, T(0), -w0, -w1, w0, T(0)};
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const T I = (i % 4 == 0) ? T(1) : T(0);
    out[i] = I - b * K[i] + c3 * (ww[i] - th2 * I);
    // R = c I + a [w]x + b w w^T  (the same a, b, c as the Jacobian code)
    out[9 + i] = k.small ? (I + K[i]) : (k.c * I + k.a * K[i] + k.b * ww[i]);
  }
  const T f = cam[6], k1 = cam[7], k2 = cam[8];
  out[18] = T(1); out[19] = T(0); out[20] = T(0);
  out[21] = k1;   out[22] = f;    out[23] = T(0);
  out[24] = k2;   out[25] = T(0); out[26] = f;
}

}  // namespace ba
