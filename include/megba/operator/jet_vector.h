// megba/operator/jet_vector.h -- JetVector<T>, the scalar type user residuals are written in
// (reference: include/operator/jet_vector.h:23-171 + ~45 CUDA kernels in src/operator/jet_vector_math_impl.cu).
//
// In the reference a JetVector owns device slabs (value[nE], grad[N][nE]) and every arithmetic operator launches a
// kernel over all edges.  Here BaseEdge::forward() runs ONCE on the host with symbolic JetVectors: every operation
// appends a node to a tape, the tape is compiled to a register bytecode (dead code removed, registers reused by
// liveness), and the library evaluates that program per observation on the GPU with forward-mode automatic
// differentiation (megba_b200/csrc/ba_program.cu) -- one fused pass instead of one launch and one slab per operator.
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace MegBA {
namespace prog {
// keep in sync with megba_b200/csrc/ba_program.cu
enum Op : uint8_t { INPUT = 0, CONST = 1, ADD = 2, SUB = 3, MUL = 4, DIV = 5, NEG = 6, SQRT = 7, SIN = 8, COS = 9, SELGT = 10, ABS = 11 };

struct Node {
  Op op;
  int a, b, c, d;  // operand nodes (INPUT: a = input number; SELGT: a > b ? c : d)
  double k;        // CONST value
};

struct Tape {
  std::vector<Node> nodes;
  int push(Op op, int a = -1, int b = -1, int c = -1, int d = -1, double k = 0.) {
    nodes.push_back(Node{op, a, b, c, d, k});
    return (int)nodes.size() - 1;
  }
};

inline Tape *&current() {
  static thread_local Tape *t = nullptr;
  return t;
}
inline Tape &tape() {
  if (!current()) throw std::runtime_error("JetVector arithmetic outside BaseEdge::forward(): no residual tape is being recorded");
  return *current();
}

// Register bytecode: instr = {op, dst, a, b, c, d} (slots), CONST: a = index into consts, INPUT: a = input number.
struct Program {
  std::vector<uint8_t> code;
  std::vector<double> consts;
  int n_slots = 0;
  int out[2] = {0, 0};
  int n_instr() const { return (int)code.size() / 6; }
};

inline Program compile(const Tape &t, int out0, int out1, int max_slots) {
  const int n = (int)t.nodes.size();
  if (out0 < 0 || out1 < 0 || out0 >= n || out1 >= n) throw std::runtime_error("forward() must return a 2x1 residual");
  // reachability (dead-code elimination) and last use
  std::vector<char> live(n, 0);
  std::vector<int> last(n, -1);
  live[out0] = live[out1] = 1;
  for (int i = n - 1; i >= 0; --i) {
    if (!live[i]) continue;
    const Node &nd = t.nodes[i];
    const int ops[4] = {nd.op == INPUT ? -1 : nd.a, nd.b, nd.c, nd.d};
    for (int o : ops)
      if (o >= 0) { live[o] = 1; if (last[o] < i) last[o] = i; }
  }
  last[out0] = last[out1] = n;  // results stay alive to the end
  Program p;
  std::vector<int> slot(n, -1), free_slots;
  int next = 0;
  auto put = [&](int v) {
    if (v < 0 || v > 255) throw std::runtime_error("residual program: operand out of range");
    p.code.push_back((uint8_t)v);
  };
  for (int i = 0; i < n; ++i) {
    if (!live[i]) continue;
    const Node &nd = t.nodes[i];
    // operands read their slots first; slots whose value dies here may be reused by the result
    const int ops[4] = {nd.op == INPUT || nd.op == CONST ? -1 : nd.a, nd.b, nd.c, nd.d};
    int os[4];
    for (int q = 0; q < 4; ++q) os[q] = ops[q] >= 0 ? slot[ops[q]] : 0;
    for (int q = 0; q < 4; ++q) {
      if (ops[q] >= 0 && last[ops[q]] == i) {
        bool dup = false;
        for (int q2 = 0; q2 < q; ++q2) dup = dup || ops[q2] == ops[q];
        if (!dup) free_slots.push_back(slot[ops[q]]);
      }
    }
    int dst;
    if (!free_slots.empty()) { dst = free_slots.back(); free_slots.pop_back(); }
    else dst = next++;
    if (next > max_slots) throw std::runtime_error("residual program needs more registers than the device interpreter has");
    slot[i] = dst;
    put(nd.op);
    put(dst);
    if (nd.op == INPUT) { put(nd.a); put(0); put(0); put(0); }
    else if (nd.op == CONST) {
      int ci = -1;
      for (std::size_t q = 0; q < p.consts.size(); ++q) if (p.consts[q] == nd.k && std::signbit(p.consts[q]) == std::signbit(nd.k)) ci = (int)q;
      if (ci < 0) { p.consts.push_back(nd.k); ci = (int)p.consts.size() - 1; }
      put(ci); put(0); put(0); put(0);
    } else { put(os[0]); put(os[1]); put(os[2]); put(os[3]); }
  }
  p.n_slots = next;
  p.out[0] = slot[out0];
  p.out[1] = slot[out1];
  return p;
}
}  // namespace prog

template <typename T>
class JetVector {
  int id_ = -1;

 public:
  JetVector() = default;
  JetVector(T c) : id_(prog::tape().push(prog::CONST, -1, -1, -1, -1, (double)c)) {}  // NOLINT: scalars promote like the reference's
  static JetVector input(int k) { JetVector j; j.id_ = prog::tape().push(prog::INPUT, k); return j; }
  static JetVector node(prog::Op op, int a, int b = -1, int c = -1, int d = -1) { JetVector j; j.id_ = prog::tape().push(op, a, b, c, d); return j; }
  int id() const {
    if (id_ < 0) throw std::runtime_error("use of an unbound JetVector");
    return id_;
  }
  bool bound() const { return id_ >= 0; }

  JetVector operator-() const { return node(prog::NEG, id()); }
  JetVector &operator+=(const JetVector &o) { return *this = *this + o; }
  JetVector &operator-=(const JetVector &o) { return *this = *this - o; }
  JetVector &operator*=(const JetVector &o) { return *this = *this * o; }
  JetVector &operator/=(const JetVector &o) { return *this = *this / o; }
  friend JetVector operator+(const JetVector &a, const JetVector &b) { return node(prog::ADD, a.id(), b.id()); }
  friend JetVector operator-(const JetVector &a, const JetVector &b) { return node(prog::SUB, a.id(), b.id()); }
  friend JetVector operator*(const JetVector &a, const JetVector &b) { return node(prog::MUL, a.id(), b.id()); }
  friend JetVector operator/(const JetVector &a, const JetVector &b) { return node(prog::DIV, a.id(), b.id()); }
  friend JetVector operator+(const JetVector &a, T b) { return a + JetVector(b); }
  friend JetVector operator+(T a, const JetVector &b) { return JetVector(a) + b; }
  friend JetVector operator-(const JetVector &a, T b) { return a - JetVector(b); }
  friend JetVector operator-(T a, const JetVector &b) { return JetVector(a) - b; }
  friend JetVector operator*(const JetVector &a, T b) { return a * JetVector(b); }
  friend JetVector operator*(T a, const JetVector &b) { return JetVector(a) * b; }
  friend JetVector operator/(const JetVector &a, T b) { return a / JetVector(b); }
  friend JetVector operator/(T a, const JetVector &b) { return JetVector(a) / b; }
};

namespace math {
template <typename T> JetVector<T> sqrt(const JetVector<T> &a) { return JetVector<T>::node(prog::SQRT, a.id()); }
template <typename T> JetVector<T> sin(const JetVector<T> &a) { return JetVector<T>::node(prog::SIN, a.id()); }
template <typename T> JetVector<T> cos(const JetVector<T> &a) { return JetVector<T>::node(prog::COS, a.id()); }
template <typename T> JetVector<T> abs(const JetVector<T> &a) { return JetVector<T>::node(prog::ABS, a.id()); }
// a > b ? x : y, evaluated per observation (value and derivatives of the chosen branch)
template <typename T>
JetVector<T> select_gt(const JetVector<T> &a, const JetVector<T> &b, const JetVector<T> &x, const JetVector<T> &y) {
  return JetVector<T>::node(prog::SELGT, a.id(), b.id(), x.id(), y.id());
}
}  // namespace math
}  // namespace MegBA
