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
    for exe in ("BAL_Double_analytical", "BAL_Float_analytical", "BAL_Double_analytical_implicit", "BAL_Double", "BAL_Float",
                "BAL_Double_implicit"):
        assert os.path.exists(os.path.join(BIN, exe))
    if HAVE_GPU:
        pytest.skip("GPU present: the no-device path cannot be exercised")
    P = make_synthetic(5, 40, 130, seed=9)
    path = str(tmp_path / "problem-5-40-pre.txt")
    write_bal(path, P)
    r = _run("BAL_Double_analytical", path, "--max_iter", "2", expect_ok=False)
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)


def _failing_rank(tmp_path):
    """--world_size 2 where not every rank gets a device (none without a GPU; only rank 0 on a 1-GPU box): the ranks
    agree on failure before the first collective, solve() throws instead of waiting in ncclCommInitRank."""
    _build()
    P = make_synthetic(5, 40, 130, seed=9)
    path = str(tmp_path / "problem-5-40-pre.txt")
    write_bal(path, P)
    r = subprocess.run([os.path.join(BIN, "BAL_Double_analytical"), "--path", path, "--max_iter", "2", "--world_size", "2"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "ba_create" in (r.stdout + r.stderr)


def test_ranks_without_devices_fail_together(tmp_path):
    if HAVE_GPU:
        pytest.skip("GPU present")
    _failing_rank(tmp_path)


@pytest.mark.gpu
def test_a_rank_that_fails_alone_does_not_hang_the_others(tmp_path):
    import torch
    if torch.cuda.device_count() >= 2:
        pytest.skip("every rank has a device on this box")
    _failing_rank(tmp_path)


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


_FACADE_PRELUDE = r'''
#include <cstdio>
#include <memory>
#include <tuple>
#include <vector>
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
'''


def _compile(tmp_path, body, name):
    cpp = tmp_path / f"{name}.cpp"
    cpp.write_text(_FACADE_PRELUDE + body)
    exe = tmp_path / name
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{ROOT}/include/megba", f"-I{ROOT}/include", "-o", str(exe),
                    str(cpp), f"-L{ROOT}/megba_b200", "-lmegba_b200", f"-Wl,-rpath,{ROOT}/megba_b200", "-pthread"],
                   check=True)
    return exe


@pytest.mark.gpu
def test_facade_vertex_api_errors(tmp_path):
    """appendVertex / getVertex / solve error behaviour of the reference (src/problem/base_problem.cpp:75-112)."""
    body = r'''
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
    exe = _compile(tmp_path, body, "api")
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "API_OK 1" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_facade_information_matrices(tmp_path):
    """Non-identity information matrices set through BaseEdge::setInformation reach ba_set_information: one LM
    iteration through the C++ facade equals the oracle's step on H = J^T Omega J (JMulInfo / makeHSchurWithInfo,
    src/edge/build_linear_system.cu:149-239)."""
    body = r'''
int main(int argc, char **argv) {
  typedef double T;
  FILE *f = std::fopen(argv[1], "r");
  int nC, nP, nE;
  if (!f || std::fscanf(f, "%d %d %d", &nC, &nP, &nE) != 3) return 2;
  MegBA::ProblemOption po; po.N = 12; po.deviceUsed = {0}; po.nItem = nE;
  MegBA::SolverOption so; so.solverOptionPCG.maxIter = 30; so.solverOptionPCG.tol = 0; so.solverOptionPCG.refuseRatio = 1e300;
  MegBA::AlgoOption ao; ao.algoOptionLM.maxIter = 1; ao.algoOptionLM.initialRegion = 1e3;
  ao.algoOptionLM.epsilon1 = 0; ao.algoOptionLM.epsilon2 = 0;
  std::unique_ptr<MegBA::BaseAlgo<T>> algo{new MegBA::LMAlgo<T>{po, ao}};
  std::unique_ptr<MegBA::BaseSolver<T>> solver{new MegBA::SchurPCGSolver<T>{po, so}};
  std::unique_ptr<MegBA::BaseLinearSystem<T>> ls{new MegBA::SchurLMLinearSystem<T>{po, std::move(solver)}};
  MegBA::BaseProblem<T> problem{po, std::move(algo), std::move(ls)};
  struct Obs { int c, p; double u, v, w[4]; };
  std::vector<Obs> obs(nE);
  for (auto &o : obs)
    if (std::fscanf(f, "%d %d %lf %lf %lf %lf %lf %lf", &o.c, &o.p, &o.u, &o.v, &o.w[0], &o.w[1], &o.w[2], &o.w[3]) != 8) return 3;
  for (int n = 0; n < nC; ++n) {
    Eigen::Matrix<T, 9, 1> x;
    for (int k = 0; k < 9; ++k) { double t; if (std::fscanf(f, "%lf", &t) != 1) return 4; x(k) = t; }
    problem.appendVertex(n, new MegBA::CameraVertex<T>());
    problem.getVertex(n).setEstimation(std::move(x));
  }
  for (int n = 0; n < nP; ++n) {
    Eigen::Matrix<T, 3, 1> x;
    for (int k = 0; k < 3; ++k) { double t; if (std::fscanf(f, "%lf", &t) != 1) return 5; x(k) = t; }
    problem.appendVertex(nC + n, new MegBA::PointVertex<T>());
    problem.getVertex(nC + n).setEstimation(std::move(x));
  }
  if (argc > 2) {  // BaseVertex::fixed through the facade
    problem.getVertex(0).fixed = true;
    problem.getVertex(nC + 3).fixed = true;
  }
  for (const auto &o : obs) {
    auto *e = new E<T>;
    e->appendVertex(&problem.getVertex(o.c));
    e->appendVertex(&problem.getVertex(nC + o.p));
    Eigen::Matrix<T, 2, 1> m; m(0) = o.u; m(1) = o.v;
    e->setMeasurement(std::move(m));
    Eigen::Matrix<T, 2, 2> w; w(0, 0) = o.w[0]; w(0, 1) = o.w[1]; w(1, 0) = o.w[2]; w(1, 1) = o.w[3];
    e->setInformation(std::move(w));
    problem.appendEdge(*e);
  }
  problem.solve();
  for (int n = 0; n < nC; ++n) for (int k = 0; k < 9; ++k) std::printf("X %.17g\n", problem.getVertex(n).getEstimation()(k));
  for (int n = 0; n < nP; ++n) for (int k = 0; k < 3; ++k) std::printf("X %.17g\n", problem.getVertex(nC + n).getEstimation()(k));
  return 0;
}
'''
    exe = _compile(tmp_path, body, "info")
    P = make_synthetic(8, 120, 500, seed=21)
    rng = np.random.default_rng(4)
    a, b, c = rng.uniform(0.5, 2.0, P.nE), rng.uniform(-0.3, 0.3, P.nE), rng.uniform(0.5, 2.0, P.nE)
    info = np.stack([a, b, b, c], axis=1)
    path = tmp_path / "problem.txt"
    with open(path, "w") as fh:
        fh.write(f"{P.nC} {P.nP} {P.nE}\n")
        for j in range(P.nE):
            fh.write(f"{P.cam_idx[j]} {P.pt_idx[j]} {P.obs[j, 0]:.17g} {P.obs[j, 1]:.17g} " + " ".join(f"{v:.17g}" for v in info[j]) + "\n")
        for v in P.cam.ravel():
            fh.write(f"{v:.17g}\n")
        for v in P.pt.ravel():
            fh.write(f"{v:.17g}\n")
    r = subprocess.run([str(exe), str(path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    x = np.array([float(l.split()[1]) for l in r.stdout.splitlines() if l.startswith("X ")])
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, np.float64)
    rr, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(rr, J, info=info)
    HppD, _ = O.process_diag(Hpp, 1e3)
    HllD, _ = O.process_diag(Hll, 1e3)
    dx, _, _, _ = O.schur_pcg(HppD, HllD, W, None, g, np.zeros(9 * P.nC + 3 * P.nP), 30, 0.0, 1e300)
    x_expected = P.x() + dx  # the step is accepted (the cost decreases on this problem)
    r1, _ = O.forward(x_expected[:9 * P.nC].reshape(-1, 9), x_expected[9 * P.nC:].reshape(-1, 3), jac=False)
    assert O.cost(r1) < O.cost(rr)
    assert np.linalg.norm(x - x_expected) <= 1e-8 * np.linalg.norm(dx)
    # and the step really depends on the information matrices
    Hpp_i, Hll_i, W_i, g_i = O.build(rr, J)
    dx_i, _, _, _ = O.schur_pcg(O.process_diag(Hpp_i, 1e3)[0], O.process_diag(Hll_i, 1e3)[0], W_i, None, g_i,
                                np.zeros(9 * P.nC + 3 * P.nP), 30, 0.0, 1e300)
    assert np.linalg.norm(dx - dx_i) > 1e-3 * np.linalg.norm(dx)

    # the same program with camera 0 and point 3 fixed (BaseVertex::fixed -> ba_set_fixed): they keep their parameters
    r = subprocess.run([str(exe), str(path), "fixed"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    xf = np.array([float(l.split()[1]) for l in r.stdout.splitlines() if l.startswith("X ")])
    x0 = P.x()
    assert np.array_equal(xf[:9], x0[:9]) and np.array_equal(xf[9 * P.nC + 9:9 * P.nC + 12], x0[9 * P.nC + 9:9 * P.nC + 12])
    assert np.abs(xf[9:18] - x0[9:18]).max() > 0  # the free cameras still move
