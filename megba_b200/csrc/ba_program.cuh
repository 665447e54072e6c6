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
//  11 ABS (a)   sign = a > 0 ? 1 : -1 as abs_JetVector_Kernel, src/operator/jet_vector_math_impl.cu:1193-1206
// library-internal (not accepted from callers; produced by ba_program_host.h):
//  12 WIDEN   derivative groups outside the operand's set become explicit zeros (the two results)
//  13 LOADF   dst = frontier value a of this observation's camera (value + camera derivatives from the per-camera table)
//  14 STOREF  frontier value b := slot[a]   (per-camera program only)
// dst may be one of the operand registers (the host-side compiler reuses registers by liveness), so every rule
// below reads its operands' k-th derivative before it writes the result's.
#pragma once
#include <stdint.h>

namespace ba {

constexpr int PROG_SLOTS = 48;  // == BA_PROGRAM_MAX_SLOTS (include/megba_b200.h)
constexpr int PROG_NIN = 14, PROG_NDIFF = 12;
constexpr int PROG_FSTRIDE = 10;  // scalars per frontier value in the per-camera table: value + 9 camera derivatives
enum ProgOp {
  P_INPUT = 0, P_CONST, P_ADD, P_SUB, P_MUL, P_DIV, P_NEG, P_SQRT, P_SIN, P_COS, P_SELGT, P_ABS,
  P_LAST_PUBLIC = P_ABS, P_WIDEN, P_LOADF, P_STOREF
};

__device__ __forceinline__ double prog_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float prog_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ void prog_sincos(double x, double &s, double &c) { sincos(x, &s, &c); }
__device__ __forceinline__ void prog_sincos(float x, float &s, float &c) { sincosf(x, &s, &c); }

// in[14] -> res[2] (+ J[2][12] when GRAD).
// Derivatives are kept in four groups of three (rotation, translation, intrinsics, point).  The host
// (ImplT::set_residual_program) propagates which groups each register can depend on and stores, per instruction,
// the group sets of the result (md) and of the operands (ma, mb) in bits 16-27 of y: a group outside md is neither
// computed nor stored, an operand group outside ma / mb reads as zero.  The sets are the same for every thread, so
// the tests on them are uniform branches.  (The rotation half of a BAL residual only ever touches group 0.)
template <typename T, bool GRAD>
__device__ __noinline__ void prog_eval(const uint2 *__restrict__ code, const T *__restrict__ consts, int n, int out0, int out1,
                                       const T *in, T *res, T *J, const T *__restrict__ tab_in, T *tab_out) {
  constexpr int N = PROG_NDIFF;
  T v[PROG_SLOTS];
  T g[GRAD ? PROG_SLOTS * N : 1];
  for (int i = 0; i < n; ++i) {
    const uint2 w = __ldg(code + i);
    const int op = w.x & 0xff, dst = (w.x >> 8) & 0xff, a = (w.x >> 16) & 0xff, b = w.x >> 24;
    const unsigned md = (w.y >> 16) & 0xf, ma = (w.y >> 20) & 0xf, mb = (w.y >> 24) & 0xf;
    T *gd = g + (GRAD ? N * dst : 0);
    const T *ga = g + (GRAD ? N * a : 0), *gb = g + (GRAD ? N * b : 0);
#define PROG_GROUPS(expr)                                             \
  if (GRAD) {                                                         \
    _Pragma("unroll") for (int q = 0; q < 4; ++q) {                   \
      if (md >> q & 1) {                                              \
        const bool ha = ma >> q & 1, hb = mb >> q & 1;                \
        (void)ha; (void)hb;                                           \
        _Pragma("unroll") for (int k = 3 * q; k < 3 * q + 3; ++k) {   \
          const T A = ha ? ga[k] : T(0), B = hb ? gb[k] : T(0);       \
          (void)A; (void)B;                                           \
          gd[k] = (expr);                                             \
        }                                                             \
      }                                                               \
    }                                                                 \
  }
    switch (op) {
      case 0: {
        PROG_GROUPS(k == a ? T(1) : T(0))
        v[dst] = in[a];
      } break;
      case 1: {
        v[dst] = __ldg(consts + a);
      } break;
      case 2: {
        const T r = v[a] + v[b];
        PROG_GROUPS(A + B)
        v[dst] = r;
      } break;
      case 3: {
        const T r = v[a] - v[b];
        PROG_GROUPS(A - B)
        v[dst] = r;
      } break;
      case 4: {
        const T va = v[a], vb = v[b];
        PROG_GROUPS(va * B + vb * A)
        v[dst] = va * vb;
      } break;
      case 5: {
        const T ib = T(1) / v[b], qv = v[a] * ib;
        PROG_GROUPS((A - qv * B) * ib)
        v[dst] = qv;
      } break;
      case 6: {
        const T r = -v[a];
        PROG_GROUPS(-A)
        v[dst] = r;
      } break;
      case 7: {
        const T s = prog_sqrt(v[a]), h = T(0.5) / s;
        PROG_GROUPS(h * A)
        v[dst] = s;
      } break;
      case 8:
      case 9: {
        T s, c;
        prog_sincos(v[a], s, c);
        const T f = op == 8 ? c : -s;
        PROG_GROUPS(f * A)
        v[dst] = op == 8 ? s : c;
      } break;
      case 10: {  // SELGT: ma / mb are the group sets of the two candidates
        const bool first = v[a] > v[b];
        const int src = first ? (int)(w.y & 0xff) : (int)((w.y >> 8) & 0xff);
        if (GRAD) {
          const T *gs = g + N * src;
          const unsigned ms = first ? ma : mb;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (md >> q & 1) {
              const bool hs = ms >> q & 1;
#pragma unroll
              for (int k = 3 * q; k < 3 * q + 3; ++k) gd[k] = hs ? gs[k] : T(0);
            }
          }
        }
        v[dst] = v[src];
      } break;
      case 11: {  // ABS
        const T va = v[a], sg = va > T(0) ? T(1) : T(-1);
        PROG_GROUPS(sg * A)
        v[dst] = sg * va;
      } break;
      case 12: {  // WIDEN
        PROG_GROUPS(A)
        v[dst] = v[a];
      } break;
      case 13: {  // LOADF: a camera-only value, evaluated once per camera by the per-camera program
        const T *src = tab_in + PROG_FSTRIDE * a;
        v[dst] = __ldg(src);
        if (GRAD) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (md >> q & 1) {
#pragma unroll
              for (int k = 3 * q; k < 3 * q + 3; ++k) gd[k] = __ldg(src + 1 + k);
            }
          }
        }
      } break;
      default: {  // 14 STOREF
        T *o = tab_out + PROG_FSTRIDE * b;
        o[0] = v[a];
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < 9; ++k) o[1 + k] = (ma >> (k / 3) & 1) ? ga[k] : T(0);
        }
      } break;
    }
#undef PROG_GROUPS
  }
  if (!res) return;  // per-camera program: the stores were the result
  res[0] = v[out0];
  res[1] = v[out1];
  if (GRAD) {
#pragma unroll
    for (int k = 0; k < N; ++k) { J[k] = g[N * out0 + k]; J[N + k] = g[N * out1 + k]; }
  }
}

// the calling convention of edge_residual_jacobian / edge_residual (ba_geometry.cuh)
// c: the observation's camera; tab: the frontier table that matches cam (Dev::prog_tab at the current parameters,
// Dev::prog_tab2 at the trial parameters)
template <typename T, bool GRAD, typename D>
__device__ __forceinline__ void prog_edge(const D &d, int c, const T *tab, const T *cam, T X0, T X1, T X2, T o0, T o1, T &r0, T &r1,
                                          T *J) {
  T in[PROG_NIN], res[2];
#pragma unroll
  for (int k = 0; k < 9; ++k) in[k] = cam[k];
  in[9] = X0; in[10] = X1; in[11] = X2; in[12] = o0; in[13] = o1;
  prog_eval<T, GRAD>(d.prog_code, d.prog_const, d.prog_n, d.prog_out0, d.prog_out1, in, res, J,
                     tab + (size_t)c * (PROG_FSTRIDE * d.prog_nf), (T *)nullptr);
  r0 = res[0];
  r1 = res[1];
}

}  // namespace ba
