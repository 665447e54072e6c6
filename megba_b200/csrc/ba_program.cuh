// ba_program.cuh -- device interpreter of a user-defined residual (ba_set_residual_program).
//
// Reference path replaced: a BaseEdge::forward() written on JetVectors (examples/BAL_Double.cpp:15-33 of the
// reference).  There every JetVector operator is its own kernel over all edges, reading and writing value[nE] and
// grad[12][nE] slabs in HBM (src/operator/jet_vector_math_impl.cu, ~45 kernels; a BAL residual is ~150 operator
// launches per linearisation).  Here the expression is a register bytecode, and ONE thread evaluates it for one
// observation with forward-mode automatic differentiation: value + 12 partial derivatives per register, in thread-
// local memory, never in HBM.  The result feeds the same Jacobian-stream kernels the analytical model uses.
//
// Instruction (device form, 8 bytes): x = op | dst << 8 | a << 16 | b << 24, y = c | d << 8.
//   0 INPUT  dst = input[a]        a: 0-8 camera, 9-11 point, 12-13 measurement (no derivative)
//   1 CONST  dst = consts[a]
//   2 ADD 3 SUB 4 MUL 5 DIV (a, b)   6 NEG 7 SQRT 8 SIN 9 COS (a)   10 SELGT  dst = slot[a] > slot[b] ? slot[c] : slot[d]
// dst may be one of the operand registers (the host-side compiler reuses registers by liveness), so every rule
// below reads its operands' k-th derivative before it writes the result's.
#pragma once
#include <stdint.h>

namespace ba {

constexpr int PROG_SLOTS = 48;  // == BA_PROGRAM_MAX_SLOTS (include/megba_b200.h)
constexpr int PROG_NIN = 14, PROG_NDIFF = 12;

__device__ __forceinline__ double prog_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float prog_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ void prog_sincos(double x, double &s, double &c) { sincos(x, &s, &c); }
__device__ __forceinline__ void prog_sincos(float x, float &s, float &c) { sincosf(x, &s, &c); }

// in[14] -> res[2] (+ J[2][12] when GRAD)
template <typename T, bool GRAD>
__device__ __noinline__ void prog_eval(const uint2 *__restrict__ code, const T *__restrict__ consts, int n, int out0, int out1,
                                       const T *in, T *res, T *J) {
  constexpr int N = PROG_NDIFF;
  T v[PROG_SLOTS];
  T g[GRAD ? PROG_SLOTS * N : 1];
  for (int i = 0; i < n; ++i) {
    const uint2 w = __ldg(code + i);
    const int op = w.x & 0xff, dst = (w.x >> 8) & 0xff, a = (w.x >> 16) & 0xff, b = w.x >> 24;
    T *gd = g + (GRAD ? N * dst : 0);
    const T *ga = g + (GRAD ? N * a : 0), *gb = g + (GRAD ? N * b : 0);
    switch (op) {
      case 0: {
        v[dst] = in[a];
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < N; ++k) gd[k] = k == a ? T(1) : T(0);
        }
      } break;
      case 1: {
        v[dst] = __ldg(consts + a);
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < N; ++k) gd[k] = T(0);
        }
      } break;
      case 2: {
        const T r = v[a] + v[b];
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < N; ++k) gd[k] = ga[k] + gb[k];
        }
        v[dst] = r;
      } break;
      case 3: {
        const T r = v[a] - v[b];
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < N; ++k) gd[k] = ga[k] - gb[k];
        }
        v[dst] = r;
      } break;
      case 4: {
        const T va = v[a], vb = v[b];
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < N; ++k) gd[k] = va * gb[k] + vb * ga[k];
        }
        v[dst] = va * vb;
      } break;
      case 5: {
        const T ib = T(1) / v[b], q = v[a] * ib;
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < N; ++k) gd[k] = (ga[k] - q * gb[k]) * ib;
        }
        v[dst] = q;
      } break;
      case 6: {
        const T r = -v[a];
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < N; ++k) gd[k] = -ga[k];
        }
        v[dst] = r;
      } break;
      case 7: {
        const T s = prog_sqrt(v[a]), h = T(0.5) / s;
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < N; ++k) gd[k] = h * ga[k];
        }
        v[dst] = s;
      } break;
      case 8:
      case 9: {
        T s, c;
        prog_sincos(v[a], s, c);
        const T f = op == 8 ? c : -s;
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < N; ++k) gd[k] = f * ga[k];
        }
        v[dst] = op == 8 ? s : c;
      } break;
      default: {  // 10 SELGT
        const int src = v[a] > v[b] ? (int)(w.y & 0xff) : (int)((w.y >> 8) & 0xff);
        if (GRAD) {
          const T *gs = g + N * src;
#pragma unroll
          for (int k = 0; k < N; ++k) gd[k] = gs[k];
        }
        v[dst] = v[src];
      } break;
    }
  }
  res[0] = v[out0];
  res[1] = v[out1];
  if (GRAD) {
#pragma unroll
    for (int k = 0; k < N; ++k) { J[k] = g[N * out0 + k]; J[N + k] = g[N * out1 + k]; }
  }
}

// the calling convention of edge_residual_jacobian / edge_residual (ba_geometry.cuh)
template <typename T, bool GRAD, typename D>
__device__ __forceinline__ void prog_edge(const D &d, const T *cam, T X0, T X1, T X2, T o0, T o1, T &r0, T &r1, T *J) {
  T in[PROG_NIN], res[2];
#pragma unroll
  for (int k = 0; k < 9; ++k) in[k] = cam[k];
  in[9] = X0; in[10] = X1; in[11] = X2; in[12] = o0; in[13] = o1;
  prog_eval<T, GRAD>(d.prog_code, d.prog_const, d.prog_n, d.prog_out0, d.prog_out1, in, res, J);
  r0 = res[0];
  r1 = res[1];
}

}  // namespace ba
