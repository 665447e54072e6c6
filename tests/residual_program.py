"""Test-side tools for user-defined residuals: trace the reference's BAL_Double edge through the C++ facade (host only,
no GPU), parse the compiled program, and evaluate it with numpy dual numbers -- an independent restatement of the
device interpreter in megba_b200/csrc/ba_program.cuh."""
import os
import subprocess

import numpy as np

from conftest import ROOT
from megba_b200.synthetic import make_synthetic, write_bal

BIN = os.path.join(ROOT, "build", "examples")
OPS = ("INPUT", "CONST", "ADD", "SUB", "MUL", "DIV", "NEG", "SQRT", "SIN", "COS", "SELGT", "ABS")


def build_examples():
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["make", "-C", os.path.join(ROOT, "megba_b200", "csrc"), "-s"], check=True, env=env)
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "-s"], check=True, env=env)


def load_program(path):
    lines = open(path).read().split("\n")
    tag, n_instr, n_const, n_slots, o0, o1 = lines[0].split()
    assert tag == "program"
    n_instr, n_const = int(n_instr), int(n_const)
    code = np.array([[int(x) for x in ln.split()] for ln in lines[1:1 + n_instr]], np.uint8).reshape(-1, 6)
    consts = np.array([float(ln) for ln in lines[1 + n_instr:1 + n_instr + n_const]], np.float64)
    return dict(code=code, consts=consts, n_slots=int(n_slots), out=(int(o0), int(o1)))


def trace_bal_edge(tmp_path):
    """The program the facade compiles from examples/BAL_Double.cpp's forward().  The driver stops at ba_create when
    there is no GPU; the residual was traced and dumped before that."""
    build_examples()
    P = make_synthetic(3, 12, 30, seed=1)
    bal = os.path.join(str(tmp_path), "problem-3-12-pre.txt")
    dump = os.path.join(str(tmp_path), "program.txt")
    write_bal(bal, P)
    subprocess.run([os.path.join(BIN, "BAL_Double"), "--path", bal, "--max_iter", "1"], capture_output=True, text=True,
                   timeout=300, env=dict(os.environ, MEGBA_PROGRAM_DUMP=dump))
    return load_program(dump)


def eval_program(prog, inputs):
    """inputs [14, n] (camera 0-8, point 9-11, measurement 12-13) -> r [n, 2], J [n, 2, 12]."""
    inputs = np.asarray(inputs, np.float64)
    n = inputs.shape[1]
    v = np.zeros((prog["n_slots"], n))
    g = np.zeros((prog["n_slots"], 12, n))
    with np.errstate(all="ignore"):
        for op, dst, a, b, c, d in prog["code"].astype(int):
            if op == 0:
                val, grad = inputs[a].copy(), np.zeros((12, n))
                if a < 12:
                    grad[a] = 1.
            elif op == 1:
                val, grad = np.full(n, prog["consts"][a]), np.zeros((12, n))
            elif op == 2:
                val, grad = v[a] + v[b], g[a] + g[b]
            elif op == 3:
                val, grad = v[a] - v[b], g[a] - g[b]
            elif op == 4:
                val, grad = v[a] * v[b], v[a] * g[b] + v[b] * g[a]
            elif op == 5:
                val = v[a] / v[b]
                grad = (g[a] - val * g[b]) / v[b]
            elif op == 6:
                val, grad = -v[a], -g[a]
            elif op == 7:
                val = np.sqrt(v[a])
                grad = g[a] * (0.5 / val)
            elif op == 8:
                val, grad = np.sin(v[a]), np.cos(v[a]) * g[a]
            elif op == 9:
                val, grad = np.cos(v[a]), -np.sin(v[a]) * g[a]
            elif op == 10:
                m = v[a] > v[b]
                val, grad = np.where(m, v[c], v[d]), np.where(m, g[c], g[d])
            elif op == 11:
                sg = np.where(v[a] > 0, 1., -1.)
                val, grad = sg * v[a], sg * g[a]
            else:
                raise ValueError(op)
            v[dst], g[dst] = val, grad
    o0, o1 = prog["out"]
    r = np.stack([v[o0], v[o1]], 1)
    J = np.stack([g[o0].T, g[o1].T], 1)
    return r, J


def edge_inputs(P):
    """[14, nE] interpreter inputs of a BAProblem."""
    return np.concatenate([P.cam[P.cam_idx].T, P.pt[P.pt_idx].T, P.obs.T], 0)


class Asm:
    """A few lines of assembler for residuals written by hand in the tests: values are numbered in order, prog() maps
    them onto registers that are reused once a value has been read for the last time."""

    def __init__(self):
        self.code, self.consts = [], []

    def _emit(self, op, a=0, b=0, c=0, d=0):
        dst = len(self.code)
        self.code.append([op, dst, a, b, c, d])
        return dst

    def inp(self, k): return self._emit(0, k)
    def const(self, x):
        self.consts.append(float(x))
        return self._emit(1, len(self.consts) - 1)
    def add(self, a, b): return self._emit(2, a, b)
    def sub(self, a, b): return self._emit(3, a, b)
    def mul(self, a, b): return self._emit(4, a, b)
    def div(self, a, b): return self._emit(5, a, b)
    def neg(self, a): return self._emit(6, a)
    def sqrt(self, a): return self._emit(7, a)
    def sin(self, a): return self._emit(8, a)
    def cos(self, a): return self._emit(9, a)
    def abs(self, a): return self._emit(11, a)
    def prog(self, o0, o1):
        nsrc = {0: 0, 1: 0, 6: 1, 7: 1, 8: 1, 9: 1, 10: 4, 11: 1}
        last = {o0: len(self.code), o1: len(self.code)}
        for i, (op, dst, *src) in enumerate(self.code):
            for v in src[:nsrc.get(op, 2)]:
                last[v] = max(last.get(v, -1), i)
        reg, free, n, out = {}, [], 0, []
        for i, (op, dst, *src) in enumerate(self.code):
            k = nsrc.get(op, 2)
            ops = [reg[v] for v in src[:k]] + src[k:]
            for v in set(src[:k]):
                if last[v] == i:
                    free.append(reg[v])
            if free:
                reg[dst] = free.pop()
            else:
                reg[dst], n = n, n + 1
            out.append([op, reg[dst], *ops])
        return dict(code=np.array(out, np.uint8), consts=np.array(self.consts), n_slots=n, out=(reg[o0], reg[o1]))



def random_program(seed, n_ops=60):
    """A random residual: a DAG over all 14 inputs with every op of the instruction set, values kept in a tame range
    (denominators 2 + cos, square roots of 1 + x^2), camera-only, point-only and mixed subexpressions."""
    rng = np.random.default_rng(seed)
    A = Asm()
    scale = [1., 1., 1., 1., 1., 1., 1e-3, 1., 1., 1., 1., 1., 1e-2, 1e-2]  # focal length and pixels brought to O(1)
    pool = [A.sin(A.mul(A.const(scale[k]), A.inp(k))) for k in range(14)]
    for _ in range(n_ops):
        op = rng.choice(["add", "sub", "mul", "neg", "sin", "cos", "abs", "sel", "div", "sqrt"])
        a, b, c, d = (pool[i] for i in rng.integers(0, len(pool), 4))
        if op in ("add", "sub", "mul"):
            v = getattr(A, op)(a, b)
            if op == "mul":
                v = A.add(v, A.const(float(rng.normal())))
        elif op in ("neg", "sin", "cos", "abs"):
            v = getattr(A, op)(a)
        elif op == "sel":
            v = A._emit(10, a, b, c, d)
        elif op == "div":
            v = A.div(a, A.add(A.const(2.), A.cos(b)))
        else:
            v = A.sqrt(A.add(A.const(1.), A.mul(a, a)))
        pool.append(A.mul(A.const(0.5), v) if op in ("add", "sub") else v)
    # the two results mix late values with one camera and one point input each, so that every block is populated
    r0 = A.add(A.add(pool[-1], pool[-3]), A.mul(pool[0], pool[9]))
    r1 = A.add(A.sub(pool[-2], pool[-4]), A.mul(pool[4], pool[10]))
    return A.prog(r0, r1)


