"""Time the S*p edge kernel for each design variant on the Venice-shaped problem, one process per variant.
Usage: python scripts/bench_variants.py 0 5 ...                      (tile layout: env BA_VAR)
       python scripts/bench_variants.py BA_LAYOUT=tile BA_ELL_RC=0,BA_PDL=0 ...   (any environment settings)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/tmp/ba_problem_cache_%s.npz" % os.environ.get("BA_WORKLOAD", "venice-1778").replace("/", "_").replace(":", "_")


def child():
    import numpy as np
    from megba_b200 import capi
    from megba_b200.synthetic import BAProblem
    if os.environ.get("BA_LIB"):  # an experimental build of the same library (e.g. build/exp/libmegba_b200_g32.so)
        capi.LIB_PATH = os.path.join(ROOT, os.environ["BA_LIB"])
    z = np.load(CACHE)
    P = BAProblem(int(z["nC"]), int(z["nP"]), int(z["nE"]), z["cam_idx"], z["pt_idx"], z["obs"], z["cam"], z["pt"])
    dtype = np.float64 if os.environ.get("BA_DTYPE", "f64") == "f64" else np.float32
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype, implicit=os.environ.get("BA_IMPLICIT") == "1")
    S.set_params(P.cam, P.pt)
    S.linearize()
    S.process_diag(1e4)
    S.solve(max_iter=10, tol=0.0, refuse_ratio=1e300)
    S.set_timing(True)
    S.reset_stats()
    for _ in range(3):
        S.solve(max_iter=40, tol=0.0, refuse_ratio=1e300)
    st = S.stats()
    a_pcg, _ = S.algorithmic_bytes()
    k = st["schur_apply_ms"] / st["schur_apply_timed"]
    print(f"{os.environ.get('BA_SPEC','-'):28s} edge {k*1e3:8.1f} us  {a_pcg/k/1e6:7.0f} GB/s   pcg loop "
          f"{st['pcg_loop_ms']/st['pcg_iters']*1e3:7.1f} us/iter   cam kernel {st['pcg_cam_ms']/max(st['pcg_cam_timed'],1)*1e3:6.1f} us   "
          f"solve {st['solve_ms']/3:7.2f} ms", flush=True)
    print(f"        check: |dx|, |x| = {S.step_norms()}", flush=True)  # must agree between variants to ~1e-12 relative
    # same three solves without per-launch events: host wall clock around the (synchronous) call
    import time
    S.set_timing(False)
    S.solve(max_iter=40, tol=0.0, refuse_ratio=1e300)
    t0 = time.perf_counter()
    for _ in range(3):
        S.solve(max_iter=40, tol=0.0, refuse_ratio=1e300)
    dt = (time.perf_counter() - t0) / 3
    print(f"        timing off: solve {dt*1e3:7.2f} ms wall  (vs {st['solve_ms']/3:7.2f} ms with per-launch events)", flush=True)


if __name__ == "__main__":
    if os.environ.get("BA_CHILD") == "1":
        child()
        sys.exit(0)
    import numpy as np
    from megba_b200.synthetic import make_named
    if not os.path.exists(CACHE):
        P = make_named(os.environ.get("BA_WORKLOAD", "venice-1778"))
        np.savez(CACHE, nC=P.nC, nP=P.nP, nE=P.nE, cam_idx=P.cam_idx, pt_idx=P.pt_idx, obs=P.obs, cam=P.cam, pt=P.pt)
    for v in sys.argv[1:] or ["0"]:
        env = dict(os.environ, BA_CHILD="1", BA_SPEC=v)
        if "=" in v:
            env.update(kv.split("=", 1) for kv in v.split(","))
        else:
            env.update(BA_VAR=v, BA_LAYOUT="tile")
        subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
