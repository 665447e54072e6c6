"""User-defined residuals, host side (no GPU): the facade traces the reference's automatic-differentiation BAL edge
(examples/BAL_Double.cpp of the reference, reproduced in examples/bal_edges.h) into a register program; the
program -- evaluated here with numpy dual numbers -- must reproduce the oracle's residual and analytical Jacobian,
including the first-order rotation branch of small-angle cameras."""
import numpy as np

from megba_b200.synthetic import make_synthetic
from oracle.oracle import Oracle
from megba_b200 import capi
from residual_program import OPS, edge_inputs, eval_program, random_program, trace_bal_edge
from util import rel_fro


def test_traced_bal_edge_matches_oracle(tmp_path):
    prog = trace_bal_edge(tmp_path)
    code = prog["code"]
    assert prog["n_slots"] <= 48 and set(code[:, 0]) <= set(range(len(OPS)))
    assert sorted(code[code[:, 0] == 0][:, 2]) == list(range(14))  # camera 9 + point 3 + measurement 2, each read once
    assert 10 in code[:, 0]  # the small-angle branch is a per-observation select, not a host-side decision
    P = make_synthetic(12, 300, 1500, seed=5)
    P.cam[3, :3] = 0.            # exactly the zero rotation
    P.cam[7, :3] *= 1e-9         # theta^2 far below the epsilon of the branch
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs)
    r0, J0 = O.forward(P.cam, P.pt)
    r, J = eval_program(prog, edge_inputs(P))
    assert np.isfinite(r).all() and np.isfinite(J).all()
    assert rel_fro(r, r0) <= 1e-13
    assert rel_fro(J, J0) <= 1e-11
    small = np.isin(P.cam_idx, [3, 7])
    assert small.any() and rel_fro(J[small], J0[small]) <= 1e-9


def test_compiler_reuses_registers_and_drops_dead_code(tmp_path):
    prog = trace_bal_edge(tmp_path)
    # far fewer registers than instructions: liveness-based reuse; every instruction's result is read later or returned
    assert prog["n_slots"] < len(prog["code"]) // 2
    code = prog["code"].astype(int)
    nsrc = {0: 0, 1: 0, 10: 4, 6: 1, 7: 1, 8: 1, 9: 1, 11: 1}
    live = set(prog["out"])
    for op, dst, *src in code[::-1]:
        assert dst in live, "an instruction whose result nobody reads survived dead-code elimination"
        live.discard(dst)
        live.update(src[:nsrc.get(op, 2)])
    assert not live


def test_generated_kernel_source_compiles_for_sm_100a(tmp_path):
    """ba_residual_program_source (host-only): the per-observation program printed as straight-line CUDA C++ and built
    by NVRTC for sm_100a -- no GPU needed.  The rotation is not in the per-observation kernel (it runs once per
    camera), derivative groups that are identically zero are not printed."""
    import pytest
    prog = trace_bal_edge(tmp_path)
    for dtype in (np.float64, np.float32):
        src, cubin = capi.residual_program_source(prog["code"], prog["consts"], prog["n_slots"], prog["out"], dtype)
        assert cubin > 1000
        assert 'extern "C" __global__' in src and "prog_fill" in src and "prog_cost" in src
        assert ("typedef double T;" in src) == (dtype == np.float64)
        body = src[src.index("prog_fill"):src.index("prog_cost")]
        assert "sincos" not in body and "tab2[" in body      # angle-axis -> rotation happens per camera
        assert body.count("// ") < len(prog["code"])          # fewer statements than the caller's program
    bad = prog["code"].copy()
    bad[5, 0] = 99
    with pytest.raises(RuntimeError, match="unknown op"):
        capi.residual_program_source(bad, prog["consts"], prog["n_slots"], prog["out"])


def test_generator_handles_every_op_in_both_precisions():
    """Random 200-instruction programs (every op, register reuse, camera-only / point-only / mixed subexpressions):
    the generated source builds with NVRTC for sm_100a in fp64 and fp32."""
    for seed in (1, 2, 3):
        prog = random_program(seed)
        assert set(range(12)) <= set(prog["code"][:, 0])  # all twelve public ops occur
        for dtype in (np.float64, np.float32):
            src, cubin = capi.residual_program_source(prog["code"], prog["consts"], prog["n_slots"], prog["out"], dtype)
            assert cubin > 1000 and "tab2[" in src
