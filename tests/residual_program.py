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
