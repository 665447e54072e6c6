// ba_program_host.h -- host side of ba_set_residual_program: validation of the caller's register program and its
// translation into the two device programs the interpreter (ba_program.cuh) runs.
//
//   per-observation program  what one thread evaluates for one observation;
//   per-camera program       the camera-only part of the residual (in a BAL residual: the whole angle-axis rotation,
//                            71 of 109 instructions), evaluated once per CAMERA into a table of frontier values
//                            (value + d/d camera parameters) that the per-observation program loads.  The reference
//                            recomputes it for every observation (and so did the first version of this interpreter);
//                            the hand-written analytical path makes the same cut with k_cam_consts.
//
// Both keep the caller's register numbers and instruction order (instructions are only dropped or replaced in place),
// so the caller's register allocation stays valid.
#pragma once
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "ba_program.cuh"

namespace ba {

struct ProgInstr {
  int op, dst, src[4];
  int nsrc;          // register operands (INPUT / CONST: 0, src[0] = input number / constant index)
  int def[4];        // defining instruction of each register operand
  unsigned dep;      // what the value depends on: bits 0-2 camera groups, 3 point, 4 measurement (SELGT: conditions too)
  unsigned groups;   // derivative groups (bits 0-3) the value carries
};

struct ProgDevice {
  std::vector<uint2> obs_code, cam_code;
  int n_frontier = 0;
};

inline int prog_nsrc(int op) { return op == P_INPUT || op == P_CONST ? 0 : op == P_SELGT ? 4 : (op >= P_NEG && op <= P_COS) || op == P_ABS ? 1 : 2; }

inline uint2 prog_word(int op, int dst, int a, int b, int c, int d, unsigned md, unsigned ma, unsigned mb) {
  uint2 w;
  w.x = (unsigned)op | (unsigned)dst << 8 | (unsigned)a << 16 | (unsigned)b << 24;
  w.y = (unsigned)c | (unsigned)d << 8 | md << 16 | ma << 20 | mb << 24;
  return w;
}

// Returns "" or the reason the program is refused.
inline std::string prog_parse(int n_instr, const uint8_t *code, int n_const, int n_slots, const int32_t *out_slot,
                              std::vector<ProgInstr> *out, int out_def[2]) {
  char msg[160];
  std::vector<int> def((size_t)n_slots, -1);
  out->clear();
  out->reserve((size_t)n_instr);
  for (int i = 0; i < n_instr; ++i) {
    const uint8_t *q = code + 6 * (size_t)i;
    ProgInstr in{};
    in.op = q[0];
    in.dst = q[1];
    if (in.op > P_LAST_PUBLIC) { snprintf(msg, sizeof msg, "instruction %d has unknown op %d", i, in.op); return msg; }
    if (in.dst >= n_slots) { snprintf(msg, sizeof msg, "instruction %d writes register %d of %d", i, in.dst, n_slots); return msg; }
    in.nsrc = prog_nsrc(in.op);
    for (int k = 0; k < 4; ++k) { in.src[k] = q[2 + k]; in.def[k] = -1; }
    if (in.op == P_INPUT && in.src[0] >= PROG_NIN) { snprintf(msg, sizeof msg, "instruction %d reads input %d of %d", i, in.src[0], PROG_NIN); return msg; }
    if (in.op == P_CONST && in.src[0] >= n_const) { snprintf(msg, sizeof msg, "instruction %d reads constant %d of %d", i, in.src[0], n_const); return msg; }
    for (int k = 0; k < in.nsrc; ++k) {
      if (in.src[k] >= n_slots || def[in.src[k]] < 0) {
        snprintf(msg, sizeof msg, "instruction %d reads register %d before it is written", i, in.src[k]);
        return msg;
      }
      in.def[k] = def[in.src[k]];
    }
    const std::vector<ProgInstr> &v = *out;
    if (in.op == P_INPUT) {
      const int a = in.src[0];
      in.dep = a < 9 ? 1u << (a / 3) : a < 12 ? 8u : 16u;
      in.groups = a < PROG_NDIFF ? 1u << (a / 3) : 0u;
    } else if (in.op == P_SELGT) {
      in.dep = v[in.def[0]].dep | v[in.def[1]].dep | v[in.def[2]].dep | v[in.def[3]].dep;
      in.groups = v[in.def[2]].groups | v[in.def[3]].groups;
    } else {
      for (int k = 0; k < in.nsrc; ++k) { in.dep |= v[in.def[k]].dep; in.groups |= v[in.def[k]].groups; }
    }
    def[in.dst] = i;
    out->push_back(in);
  }
  for (int k = 0; k < 2; ++k) {
    if (out_slot[k] < 0 || out_slot[k] >= n_slots || def[out_slot[k]] < 0) {
      snprintf(msg, sizeof msg, "result %d is register %d, which no instruction writes", k, out_slot[k]);
      return msg;
    }
    out_def[k] = def[out_slot[k]];
  }
  return "";
}

// hoist = false: one per-observation program, nothing per camera.
inline ProgDevice prog_lower(const std::vector<ProgInstr> &P, const int out_def[2], const int32_t *out_slot, bool hoist) {
  const int n = (int)P.size();
  std::vector<char> cam_only(n, 0), needed(n, 0);
  std::vector<int> frontier(n, -1);
  for (int i = 0; i < n; ++i) cam_only[i] = hoist && (P[i].dep & 0x18u) == 0 && P[i].op != P_INPUT && P[i].op != P_CONST;
  // (1) what the per-observation program needs: walk back from the results; a hoisted value that is read becomes a
  //     frontier value and ends the walk there
  ProgDevice D;
  {
    std::vector<int> stack = {out_def[0], out_def[1]};
    while (!stack.empty()) {
      const int i = stack.back();
      stack.pop_back();
      if (needed[i]) continue;
      needed[i] = 1;
      if (cam_only[i]) { frontier[i] = 0; continue; }
      for (int k = 0; k < P[i].nsrc; ++k) stack.push_back(P[i].def[k]);
    }
    for (int i = 0; i < n; ++i)
      if (frontier[i] == 0) frontier[i] = D.n_frontier++;
    if (D.n_frontier > 255) return prog_lower(P, out_def, out_slot, false);
  }
  auto word = [&](const ProgInstr &in) {
    unsigned ma = 0, mb = 0;
    if (in.op == P_SELGT) { ma = P[in.def[2]].groups; mb = P[in.def[3]].groups; }
    else if (in.nsrc >= 1) { ma = P[in.def[0]].groups; mb = in.nsrc == 2 ? P[in.def[1]].groups : 0u; }
    return prog_word(in.op, in.dst, in.src[0], in.src[1], in.src[2], in.src[3], in.groups, ma, mb);
  };
  for (int i = 0; i < n; ++i) {
    if (!needed[i]) continue;
    if (cam_only[i]) D.obs_code.push_back(prog_word(P_LOADF, P[i].dst, frontier[i], 0, 0, 0, P[i].groups, 0, 0));
    else D.obs_code.push_back(word(P[i]));
  }
  unsigned full[2] = {P[out_def[0]].groups, P[out_def[1]].groups};
  for (int k = 0; k < 2; ++k) {  // the Jacobian rows are read in full: make the untouched groups of the results zeros
    if (k == 1 && out_slot[1] == out_slot[0]) break;
    D.obs_code.push_back(prog_word(P_WIDEN, out_slot[k], out_slot[k], 0, 0, 0, 0xf, full[k], 0));
  }
  // (2) the per-camera program: everything the frontier values need, each followed by its store
  if (D.n_frontier > 0) {
    std::vector<char> cam_needed(n, 0);
    std::vector<int> stack;
    for (int i = 0; i < n; ++i)
      if (frontier[i] >= 0) stack.push_back(i);
    while (!stack.empty()) {
      const int i = stack.back();
      stack.pop_back();
      if (cam_needed[i]) continue;
      cam_needed[i] = 1;
      for (int k = 0; k < P[i].nsrc; ++k) stack.push_back(P[i].def[k]);
    }
    for (int i = 0; i < n; ++i) {
      if (!cam_needed[i]) continue;
      D.cam_code.push_back(word(P[i]));
      if (frontier[i] >= 0) D.cam_code.push_back(prog_word(P_STOREF, 0, P[i].dst, frontier[i], 0, 0, 0, P[i].groups, 0));
    }
  }
  return D;
}

}  // namespace ba
