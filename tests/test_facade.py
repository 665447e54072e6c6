"""The C++ facade (include/megba/, MegBA's Problem / Vertex / Edge class names) and the BAL example drivers:
build from the BAL text file, solve, and compare with the oracle run on the same file."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import HAVE_GPU, ROOT
from megba_b200.synthetic import make_synthetic, read_bal, write_bal
from oracle.oracle import Oracle

BIN = os.path.join(ROOT, "build", "examples")


def _build():
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["make", "-C", os.path.join(ROOT, "megba_b200", "csrc"), "-s"], check=True, env=env)
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "-s"], check=True, env=env)


def _run(exe, path, *flags, expect_ok=True):
    cmd = [os.path.join(BIN, exe), "--path", path, *flags]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    if expect_ok:
        assert r.returncode == 0, r.stdout + r.stderr
    return r


def test_examples_build_and_fail_loudly_without_gpu(tmp_path):
    _build()
    for exe in ("BAL_Double_analytical", "BAL_Float_analytical", "BAL_Double_analytical_implicit"):
        assert os.path.exists(os.path.join(BIN, exe))
    if HAVE_GPU:
        pytest.skip("GPU present: the no-device path cannot be exercised")
    P = make_synthetic(5, 40, 130, seed=9)
    path = str(tmp_path / "problem-5-40-pre.txt")
    write_bal(path, P)
    r = _run("BAL_Double_analytical", path, "--max_iter", "2", expect_ok=False)
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("exe,dtype,implicit", [("BAL_Double_analytical", np.float64, False),
                                                ("BAL_Double_analytical_implicit", np.float64, True),
                                                ("BAL_Float_analytical", np.float32, False)])
def test_bal_example_matches_oracle(tmp_path, exe, dtype, implicit):
    _build()
    P = make_synthetic(20, 900, 4000, seed=13)
    path = str(tmp_path / "problem-20-900-pre.txt")
    write_bal(path, P)
    r = _run(exe, path, "--max_iter", "4", "--solver_max_iter", "40", "--solver_tol", "1e-6", "--tau", "1e4",
             "--epsilon1", "1e-3", "--epsilon2", "1e-12")
    out = r.stdout
    assert "Start with error:" in out and "Finished" in out
    m = re.search(r"RESULT iterations (\d+) final_cost (\S+) camera_sum (\S+)", out)
    assert m, out
    Q = read_bal(path)  # what the driver parsed
    O = Oracle(Q.nC, Q.nP, Q.cam_idx, Q.pt_idx, Q.obs, dtype)
    x, k, tr = O.lm_solve(Q.x(), implicit=implicit, lm_max_iter=4, initial_region=1e4, eps1=1e-3, eps2=1e-12,
                          pcg_max_iter=40, pcg_tol=1e-6, pcg_refuse=1.0)
    tol = 1e-8 if dtype == np.float64 else 2e-3
    assert int(m.group(1)) == len(tr) - 1
    assert abs(float(m.group(2)) - tr[-1]["cost"]) <= tol * tr[-1]["cost"]
    assert abs(float(m.group(3)) - x[:9 * Q.nC].sum()) <= max(tol, 1e-9) * abs(x[:9 * Q.nC].sum()) * 10
    # console trace format of lm_algo.cu:149-150,190-191
    assert re.search(r"Iter 1 error: \S+, log error: \S+, elapsed \d+ ms", out)


@pytest.mark.gpu
def test_facade_information_and_vertex_api(tmp_path):
    """Non-identity information matrices go through BaseEdge::setInformation -> ba_set_information."""
    src = r'''
#include <cstdio>
#include <memory>
#include "algo/lm_algo.h"
#include "edge/base_edge.h"
#include "geo/geo.cuh"
#include "linear_system/schur_LM_linear_system.h"
#include "problem/base_problem.h"
#include "solver/schur_pcg_solver.h"
#include "vertex/base_vertex.h"
template <typename T> struct E : MegBA::BaseEdge<T> {
  MegBA::JVD<T> forward() override {
    using M = Eigen::Map<const MegBA::geo::JVD<T>>;
    const auto &V = this->getVertices();
    M a{&V[0].getEstimation()(0, 0), 3, 1}, t{&V[0].getEstimation()(3, 0), 3, 1}, k{&V[0].getEstimation()(6, 0), 3, 1};
    return MegBA::geo::AnalyticalDerivativesKernelMatrix(a, t, k, V[1].getEstimation(), this->getMeasurement());
  }
};
int main() {
  typedef double T;
  MegBA::ProblemOption po; po.N = 12; po.deviceUsed = {0};
  MegBA::SolverOption so; MegBA::AlgoOption ao; ao.algoOptionLM.maxIter = 1;
  std::unique_ptr<MegBA::BaseAlgo<T>> algo{new MegBA::LMAlgo<T>{po, ao}};
  std::unique_ptr<MegBA::BaseSolver<T>> solver{new MegBA::SchurPCGSolver<T>{po, so}};
  std::unique_ptr<MegBA::BaseLinearSystem<T>> ls{new MegBA::SchurLMLinearSystem<T>{po, std::move(solver)}};
  MegBA::BaseProblem<T> problem{po, std::move(algo), std::move(ls)};
  bool threw = false;
  problem.appendVertex(0, new MegBA::CameraVertex<T>());
  try { problem.appendVertex(0, new MegBA::CameraVertex<T>()); } catch (const std::runtime_error &) { threw = true; }
  try { problem.getVertex(42); threw = false; } catch (const std::runtime_error &) {}
  try { problem.solve(); threw = false; } catch (const std::runtime_error &) {}
  std::printf("API_OK %d\n", (int)threw);
  return threw ? 0 : 1;
}
'''
    cpp = tmp_path / "t.cpp"
    cpp.write_text(src)
    exe = tmp_path / "t"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{ROOT}/include/megba", f"-I{ROOT}/include", "-o", str(exe),
                    str(cpp), f"-L{ROOT}/megba_b200", "-lmegba_b200", f"-Wl,-rpath,{ROOT}/megba_b200", "-pthread"],
                   check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "API_OK 1" in r.stdout, r.stdout + r.stderr
