// ba_program_jit.h -- a user-defined residual as a kernel of its own.
//
// The interpreter (ba_program.cuh) keeps value + 12 partials per program register in thread-local MEMORY, because
// register numbers that are data cannot name hardware registers: 4.8 ms per Venice linearisation against 0.38 ms of
// the hand-written analytical kernels.  Here the per-observation program (ba_program_host.h, after the camera-only
// part has been split off) is printed as straight-line CUDA C++ -- one scalar variable per register and partial,
// derivative groups that cannot be non-zero not printed at all -- and compiled at ba_set_residual_program time with
// NVRTC for sm_100a.  ptxas then does what it does for the hand-written kernels: everything lives in registers.
// This is the fused kernel SURVEY section 8f-3 asks for in place of the reference's one launch and one
// (12+1) x nE slab per JetVector operator (src/operator/jet_vector_math_impl.cu).
//
// Two kernels are generated:
//   prog_fill  the point side of the linearisation (what k_lin_ell does for the built-in residual): r and J of every
//              observation -> Jacobian stream, residual array, camera-major copy of the camera part for k_lin_cm<PROG>;
//              Hll, g_l and the cost as thread-local sums.
//   prog_cost  sum of squared residuals at the trial point (cam + dx_c, pt + dx_l), one partial per warp.
// NVRTC needs no GPU: ba_residual_program_source (host-only) returns the source and compiles it, which is how the
// CPU test suite checks the generator.
#pragma once
#include <nvrtc.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "ba_program_host.h"

namespace ba {

struct JitWord { int op, dst, a, b, c, d; unsigned md, ma, mb; };
inline JitWord jit_decode(uint2 w) {
  return JitWord{(int)(w.x & 0xff), (int)((w.x >> 8) & 0xff), (int)((w.x >> 16) & 0xff), (int)(w.x >> 24),
                 (int)(w.y & 0xff), (int)((w.y >> 8) & 0xff), (w.y >> 16) & 0xf, (w.y >> 20) & 0xf, (w.y >> 24) & 0xf};
}

// the per-observation program as statements on v<slot> / g<slot>_<k>; grad = false: values only
inline std::string jit_body(const std::vector<uint2> &code, const std::vector<double> &consts, bool grad, bool f64) {
  std::string s;
  char buf[256];
  auto V = [](int r) { return "v" + std::to_string(r); };
  auto G = [](int r, int k) { return "g" + std::to_string(r) + "_" + std::to_string(k); };
  auto has = [](unsigned m, int k) { return (m >> (k / 3) & 1u) != 0; };
  auto lit = [&](double x) { snprintf(buf, sizeof buf, "(T)%a", x); return std::string(buf); };
  int n = 0;
  for (uint2 word : code) {
    const JitWord w = jit_decode(word);
    snprintf(buf, sizeof buf, "  { // %d: op %d\n", n++, w.op);
    s += buf;
    std::vector<std::string> ng(12);  // new partials of the result (groups in md), "" = not produced
    std::string nv;
    switch (w.op) {
      case P_INPUT:
        nv = "in" + std::to_string(w.a);
        for (int k = 0; k < 12; ++k) if (has(w.md, k)) ng[k] = k == w.a ? "T(1)" : "T(0)";
        break;
      case P_CONST:
        nv = lit(consts[w.a]);
        break;
      case P_ADD:
      case P_SUB: {
        const char *o = w.op == P_ADD ? " + " : " - ";
        s += "    const T a_ = " + V(w.a) + ", b_ = " + V(w.b) + ";\n";
        nv = std::string("a_") + o + "b_";
        for (int k = 0; k < 12; ++k) {
          if (!has(w.md, k)) continue;
          const bool ha = has(w.ma, k), hb = has(w.mb, k);
          ng[k] = ha && hb ? G(w.a, k) + o + G(w.b, k) : ha ? G(w.a, k) : hb ? (w.op == P_ADD ? G(w.b, k) : "-" + G(w.b, k)) : "T(0)";
        }
      } break;
      case P_MUL:
        s += "    const T a_ = " + V(w.a) + ", b_ = " + V(w.b) + ";\n";
        nv = "a_ * b_";
        for (int k = 0; k < 12; ++k) {
          if (!has(w.md, k)) continue;
          const bool ha = has(w.ma, k), hb = has(w.mb, k);
          ng[k] = ha && hb ? "a_ * " + G(w.b, k) + " + b_ * " + G(w.a, k) : ha ? "b_ * " + G(w.a, k) : hb ? "a_ * " + G(w.b, k) : "T(0)";
        }
        break;
      case P_DIV:
        s += "    const T ib_ = T(1) / " + V(w.b) + ", q_ = " + V(w.a) + " * ib_;\n";
        nv = "q_";
        for (int k = 0; k < 12; ++k) {
          if (!has(w.md, k)) continue;
          const bool ha = has(w.ma, k), hb = has(w.mb, k);
          ng[k] = ha && hb ? "(" + G(w.a, k) + " - q_ * " + G(w.b, k) + ") * ib_" : ha ? G(w.a, k) + " * ib_"
                  : hb ? "-(q_ * " + G(w.b, k) + ") * ib_" : "T(0)";
        }
        break;
      case P_NEG:
        nv = "-" + V(w.a);
        for (int k = 0; k < 12; ++k) if (has(w.md, k)) ng[k] = has(w.ma, k) ? "-" + G(w.a, k) : "T(0)";
        break;
      case P_SQRT:
        s += std::string("    const T s_ = ") + (f64 ? "sqrt(" : "sqrtf(") + V(w.a) + "), h_ = T(0.5) / s_;\n";
        nv = "s_";
        for (int k = 0; k < 12; ++k) if (has(w.md, k)) ng[k] = has(w.ma, k) ? "h_ * " + G(w.a, k) : "T(0)";
        break;
      case P_SIN:
      case P_COS:
        s += std::string("    T sn_, cs_; ") + (f64 ? "sincos(" : "sincosf(") + V(w.a) + ", &sn_, &cs_);\n";
        s += w.op == P_SIN ? "    const T f_ = cs_;\n" : "    const T f_ = -sn_;\n";
        nv = w.op == P_SIN ? "sn_" : "cs_";
        for (int k = 0; k < 12; ++k) if (has(w.md, k)) ng[k] = has(w.ma, k) ? "f_ * " + G(w.a, k) : "T(0)";
        break;
      case P_SELGT:
        s += "    const bool c_ = " + V(w.a) + " > " + V(w.b) + ";\n";
        nv = "c_ ? " + V(w.c) + " : " + V(w.d);
        for (int k = 0; k < 12; ++k)
          if (has(w.md, k)) ng[k] = "c_ ? " + (has(w.ma, k) ? G(w.c, k) : std::string("T(0)")) + " : " + (has(w.mb, k) ? G(w.d, k) : std::string("T(0)"));
        break;
      case P_ABS:
        s += "    const T sg_ = " + V(w.a) + " > T(0) ? T(1) : T(-1);\n";
        nv = "sg_ * " + V(w.a);
        for (int k = 0; k < 12; ++k) if (has(w.md, k)) ng[k] = has(w.ma, k) ? "sg_ * " + G(w.a, k) : "T(0)";
        break;
      case P_WIDEN:
        nv = V(w.a);
        for (int k = 0; k < 12; ++k) if (!has(w.ma, k)) ng[k] = "T(0)";  // the groups it had stay as they are
        break;
      case P_LOADF: {  // value + camera partials of frontier value a: 8/16-byte loads of the scalars that are needed
        bool need[PROG_FSTRIDE] = {true};
        for (int k = 0; k < 9; ++k) need[1 + k] = grad && has(w.md, k);
        for (int q = 0; q < PROG_FSTRIDE / 2; ++q)
          if (need[2 * q] || need[2 * q + 1]) {
            snprintf(buf, sizeof buf, "    const T2 t%d_ = tab2[%d];\n", q, PROG_FSTRIDE / 2 * w.a + q);
            s += buf;
          }
        auto el = [&](int i) { snprintf(buf, sizeof buf, "t%d_.%c", i / 2, i % 2 ? 'y' : 'x'); return std::string(buf); };
        nv = el(0);
        for (int k = 0; k < 9; ++k) if (has(w.md, k)) ng[k] = el(1 + k);
      } break;
      default:
        break;
    }
    // every operand has been read into the temporaries below before the result overwrites a register it may share
    s += "    const T nv_ = " + nv + ";\n";
    if (grad) {
      for (int k = 0; k < 12; ++k)
        if (!ng[k].empty()) s += "    const T ng" + std::to_string(k) + "_ = " + ng[k] + ";\n";
      for (int k = 0; k < 12; ++k)
        if (!ng[k].empty()) s += "    " + G(w.dst, k) + " = ng" + std::to_string(k) + "_;\n";
    }
    s += "    " + V(w.dst) + " = nv_;\n  }\n";
  }
  return s;
}

inline std::string jit_source(const ProgDevice &D, const std::vector<double> &consts, int n_slots, const int32_t *out_slot,
                              bool f64) {
  std::string s;
  s += "// generated by megba_b200 from a residual program (ba_set_residual_program)\n";
  s += f64 ? "typedef double T;\ntypedef double2 T2;\n" : "typedef float T;\ntypedef float2 T2;\n";
  s += R"(
__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#define DECLARE_REGS \
)";
  for (int r = 0; r < n_slots; ++r) {
    s += "  T v" + std::to_string(r) + " = T(0);";
    for (int k = 0; k < 12; ++k) s += " T g" + std::to_string(r) + "_" + std::to_string(k) + " = T(0);";
    s += " \\\n";
  }
  s += "  (void)0\n";
  const std::string o0 = std::to_string(out_slot[0]), o1 = std::to_string(out_slot[1]);
  // ---- prog_fill
  s += R"(
template <typename X> __device__ __forceinline__ X warp_sum_t(X v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// One warp per slice (lane = point) or per heavy point (lane = observation): the whole point side of the linearisation,
// as k_lin_ell (ba_ell.cu) does it for the built-in residual -- r, J -> Jacobian stream, residual array, camera-major
// copy of the camera part; Hll, g_l and the cost as thread-local sums (fixed vertices and information matrices included).
extern "C" __global__ void __launch_bounds__(128) prog_fill(const int4 *__restrict__ wdesc, int nW, const int *__restrict__ r_cam,
    const T2 *__restrict__ r_obs, const T *__restrict__ cam16, const T *__restrict__ pt, const T *__restrict__ tab, int tab_stride,
    T2 *__restrict__ WJ, T2 *__restrict__ r_slots, const int *__restrict__ slot_cm, T2 *__restrict__ cm_jr,
    const T *__restrict__ r_info, const unsigned char *__restrict__ cam_fixed, const unsigned char *__restrict__ pt_fixed,
    T *__restrict__ Hll, T *__restrict__ gl, double *__restrict__ partials) {
  const int lane = threadIdx.x & 31, w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nW) return;
  const int4 dv = wdesc[w];
  const int row0 = dv.x, D = dv.y, pt0 = dv.z, heavy = dv.w & 0xff, np = (dv.w >> 8) & 0xff;
  const int p = heavy ? pt0 : pt0 + lane;
  const bool pvalid = heavy || lane < np;
  T in9 = T(0), in10 = T(0), in11 = T(0);
  if (pvalid) { in9 = pt[3 * (size_t)p]; in10 = pt[3 * (size_t)p + 1]; in11 = pt[3 * (size_t)p + 2]; }
  const bool pfixed = pvalid && pt_fixed && pt_fixed[p];
  T hl[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) hl[k] = T(0);
  double cost = 0.;
  for (int r = 0; r < D; ++r) {
    const size_t row = (size_t)row0 + r, slot = row * 32 + lane;
    const int c = r_cam[slot];
    if (c < 0) continue;
    const T2 *cp = reinterpret_cast<const T2 *>(cam16 + 16 * (size_t)c);  // 16-scalar aligned rows, the camera in the first 9
    const T2 c01 = cp[0], c23 = cp[1], c45 = cp[2], c67 = cp[3], c8 = cp[4];
    const T in0 = c01.x, in1 = c01.y, in2 = c23.x, in3 = c23.y, in4 = c45.x, in5 = c45.y, in6 = c67.x, in7 = c67.y, in8 = c8.x;
    const T2 ob = r_obs[slot];
    const T in12 = ob.x, in13 = ob.y;
    const T2 *tab2 = reinterpret_cast<const T2 *>(tab + (size_t)c * tab_stride);
    (void)in0; (void)in1; (void)in2; (void)in3; (void)in4; (void)in5; (void)in6; (void)in7; (void)in8; (void)in12; (void)in13; (void)tab2;
    DECLARE_REGS;
)";
  s += jit_body(D.obs_code, consts, true, f64);
  s += "    T J[24];\n";
  for (int k = 0; k < 24; ++k) s += "    J[" + std::to_string(k) + "] = g" + (k < 12 ? o0 : o1) + "_" + std::to_string(k % 12) + ";\n";
  s += "    const T r0 = v" + o0 + ", r1 = v" + o1 + ";\n";
  s += R"(
    {  // camera-major copy for the camera side (k_lin_cm, which drops the columns of a fixed camera itself)
      T2 *cm = cm_jr + 10 * (size_t)slot_cm[slot];
#pragma unroll
      for (int j = 0; j < 4; ++j) { T2 x; x.x = J[2 * j]; x.y = J[2 * j + 1]; cm[j] = x; }
      { T2 x; x.x = J[8]; x.y = J[12]; cm[4] = x; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { T2 x; x.x = J[13 + 2 * j]; x.y = J[14 + 2 * j]; cm[5 + j] = x; }
      { T2 x; x.x = r0; x.y = r1; cm[9] = x; }
    }
    if (pfixed) {  // a fixed vertex has no columns
#pragma unroll
      for (int q = 0; q < 3; ++q) J[9 + q] = J[21 + q] = T(0);
    }
    if (cam_fixed && cam_fixed[c]) {
#pragma unroll
      for (int q = 0; q < 9; ++q) J[q] = J[12 + q] = T(0);
    }
    T2 *dst = WJ + row * (size_t)(12 * 32) + lane;
#pragma unroll
    for (int j = 0; j < 12; ++j) { T2 x; x.x = J[2 * j]; x.y = J[2 * j + 1]; dst[j * 32] = x; }
    { T2 x; x.x = r0; x.y = r1; r_slots[slot] = x; }
    cost += (double)(r0 * r0) + (double)(r1 * r1);
    T Jp[6], L[6];
#pragma unroll
    for (int q = 0; q < 3; ++q) { Jp[q] = J[9 + q]; Jp[3 + q] = J[21 + q]; }
    if (r_info) {  // Omega-weighted rows (JMulInfo)
      const T om0 = r_info[4 * slot], om1 = r_info[4 * slot + 1], om2 = r_info[4 * slot + 2], om3 = r_info[4 * slot + 3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        L[q] = Jp[q] * om0 + Jp[3 + q] * om2;
        L[3 + q] = Jp[q] * om1 + Jp[3 + q] * om3;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 6; ++q) L[q] = Jp[q];
    }
#pragma unroll
    for (int y = 0; y < 3; ++y) {
#pragma unroll
      for (int q = 0; q < 3; ++q) hl[3 * y + q] += L[y] * Jp[q] + L[3 + y] * Jp[3 + q];
      hl[9 + y] -= L[y] * r0 + L[3 + y] * r1;
    }
  }
  if (heavy) {
#pragma unroll
    for (int k = 0; k < 12; ++k) hl[k] = warp_sum_t<T>(hl[k]);
  }
  if (pvalid && (!heavy || lane == 0)) {
    if (pfixed) {  // identity block, zero gradient: dx_l = 0
#pragma unroll
      for (int k = 0; k < 9; ++k) hl[k] = (k % 4 == 0) ? T(1) : T(0);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) Hll[9 * (size_t)p + k] = hl[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) gl[3 * (size_t)p + k] = hl[9 + k];
  }
  cost = warp_sum(cost);
  if (lane == 0) partials[w] = cost;
}
)";
  // ---- prog_cost
  s += R"(
extern "C" __global__ void __launch_bounds__(128) prog_cost(const int4 *__restrict__ wdesc, int nW, const int *__restrict__ r_cam,
    const T2 *__restrict__ r_obs, const T *__restrict__ cam16, const T *__restrict__ pt,
    const T *__restrict__ dxl, const T *__restrict__ tab, int tab_stride, double *__restrict__ partials, int partial_stride) {
  const int lane = threadIdx.x & 31, w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nW) return;
  const int4 dv = wdesc[w];
  const int row0 = dv.x, D = dv.y, pt0 = dv.z, heavy = dv.w & 0xff, np = (dv.w >> 8) & 0xff;
  const int p = heavy ? pt0 : pt0 + lane;
  const bool pvalid = heavy || lane < np;
  T in9 = T(0), in10 = T(0), in11 = T(0);
  if (pvalid) {
    in9 = pt[3 * (size_t)p] + dxl[3 * (size_t)p]; in10 = pt[3 * (size_t)p + 1] + dxl[3 * (size_t)p + 1];
    in11 = pt[3 * (size_t)p + 2] + dxl[3 * (size_t)p + 2];
  }
  double cost = 0.;
  for (int r = 0; r < D; ++r) {
    const size_t slot = ((size_t)row0 + r) * 32 + lane;
    const int c = r_cam[slot];
    if (c < 0) continue;
    const T2 *cp = reinterpret_cast<const T2 *>(cam16 + 16 * (size_t)c);  // 16-scalar aligned rows, the camera in the first 9
    const T2 c01 = cp[0], c23 = cp[1], c45 = cp[2], c67 = cp[3], c8 = cp[4];
    const T in0 = c01.x, in1 = c01.y, in2 = c23.x, in3 = c23.y, in4 = c45.x, in5 = c45.y, in6 = c67.x, in7 = c67.y, in8 = c8.x;
    const T2 ob = r_obs[slot];
    const T in12 = ob.x, in13 = ob.y;
    const T2 *tab2 = reinterpret_cast<const T2 *>(tab + (size_t)c * tab_stride);
    (void)in0; (void)in1; (void)in2; (void)in3; (void)in4; (void)in5; (void)in6; (void)in7; (void)in8; (void)in12; (void)in13; (void)tab2;
    DECLARE_REGS;
)";
  s += jit_body(D.obs_code, consts, false, f64);
  s += "    cost += (double)(v" + o0 + " * v" + o0 + ") + (double)(v" + o1 + " * v" + o1 + ");\n  }\n";
  s += "  cost = warp_sum(cost);\n  if (lane == 0) partials[(size_t)w * partial_stride] = cost;\n}\n";
  return s;
}

// NVRTC: source -> cubin for sm_100a.  Returns "" or the reason (with the compiler log).
inline std::string jit_compile(const std::string &src, std::vector<char> *cubin) {
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, src.c_str(), "residual_program.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) return "nvrtcCreateProgram failed";
  const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--fmad=true"};
  const nvrtcResult rc = nvrtcCompileProgram(prog, 4, opts);
  std::string why;
  if (rc != NVRTC_SUCCESS) {
    size_t n = 0;
    nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) nvrtcGetProgramLog(prog, &log[0]);
    why = std::string("NVRTC: ") + nvrtcGetErrorString(rc) + "\n" + log;
  } else {
    size_t n = 0;
    if (nvrtcGetCUBINSize(prog, &n) != NVRTC_SUCCESS || n == 0) why = "NVRTC produced no cubin";
    else {
      cubin->resize(n);
      nvrtcGetCUBIN(prog, cubin->data());
    }
  }
  nvrtcDestroyProgram(&prog);
  return why;
}

}  // namespace ba
