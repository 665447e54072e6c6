#!/usr/bin/env python
"""bench.py -- LM iterations/s and PCG iterations/s on the Venice-1778-shaped problem (fp64).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, all host threads)

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): synthetic BAL problem
with Venice's exact shape (1778 cameras, 993923 points, 5001946 observations; no BAL file and no network
on the box), fp64, explicit Schur, fixed-work mode of SURVEY.md section 8d: every LM iteration runs exactly
`--pcg-iters` (default 50) PCG iterations (tol 0, refuse ratio inf), LM stop tests disabled.

A *step* is one Levenberg-Marquardt iteration (lm_algo.cu:163-221): damping, Schur-PCG solve, step norms,
trial cost + gain ratio, and -- when accepted -- re-linearisation.
  value : K iterations inside ONE ba_lm_solve call, everything resident in HBM (device-event timed).
  e2e   : K calls of the reference-facing flow problem.solve() with max_iter = 1 on HOST buffers:
          H2D of the parameter vector (pinned), initial linearisation, one LM iteration, D2H of the
          result (pinned) -- BaseProblem::solve = buildIndex/H2D -> algo.solve -> writeBack.
Multi-GPU: one process per GPU; points (hence observations) are sharded, cameras replicated; per PCG
iteration one NCCL allreduce of the 9*nC camera vector.  Total work is fixed => "strong" scaling.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def physical_cores():
    """Physical cores this process may run on (SMT siblings counted once); the CPU arms use exactly this many threads."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except AttributeError:
        allowed = list(range(os.cpu_count() or 1))
    seen = set()
    for c in allowed:
        base = f"/sys/devices/system/cpu/cpu{c}/topology/"
        try:
            with open(base + "core_id") as f1, open(base + "physical_package_id") as f2:
                seen.add((f2.read().strip(), f1.read().strip()))
        except OSError:
            seen.add(("?", str(c)))
    return max(1, len(seen))


# The CPU arms (oracle port) run on all physical cores with threads pinned, whatever the launcher exported
# (torch.distributed.run sets OMP_NUM_THREADS=1).  libgomp reads these once, when the first OpenMP library loads, and
# OMP_PROC_BIND also pins the process's MAIN thread -- so they are set only in a process that runs nothing but a CPU arm
# (`--impl reference`, or the child process the GPU arm spawns for its cpu_baseline).  Setting them in the GPU ranks
# would pin every rank's launching thread onto the same core.
CPU_THREADS = physical_cores()


def pin_cpu_arm_env(env):
    env["OMP_NUM_THREADS"] = str(CPU_THREADS)
    env["OMP_PROC_BIND"] = "close"
    env["OMP_PLACES"] = "cores"
    env.setdefault("OMP_DYNAMIC", "false")
    return env


if "--impl" in sys.argv and "reference" in sys.argv:
    pin_cpu_arm_env(os.environ)

import numpy as np  # noqa: E402

WORKLOAD = "venice-1778"
LM_REGION = 1e4  # README's --tau for Venice (README.md:58)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pcg-iters", type=int, default=50)
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--implicit", action="store_true")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layout", default="auto", choices=["auto", "ell", "tile"],
                    help="auto = the library default (ELL layout, factored operator stream); tile = round-1 W/J stream kernels")
    ap.add_argument("--no-readme-mode", action="store_true", help="skip the README-parameter run reported beside the fixed-work line")
    ap.add_argument("--cpu-arm-only", type=float, default=0.0, metavar="BUDGET_S",
                    help="(internal) --impl reference: time the CPU arm for the in-process cpu_baseline of the GPU arm and print its measurement")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled DURING the timed regions.  NVML in a thread of this process
    (the library nvidia-smi itself reads; a sample every 20 ms (BENCH_CLOCK_MS): the 30 ms + 45 ms timed regions of an
    8-GPU run still get a few, and four 4-GPU runs showed no effect on the timings at 50 ms and one slower camera/exchange
    kernel in two runs at 5 ms), `nvidia-smi -lms 25` in a subprocess when the NVML binding is not importable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReasons bits (nvml.h)
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.thread = None
        self.samples = []   # (sm_mhz, reasons_bitmask)
        self.max_mhz = None
        self.running = False
        self.period = float(os.environ.get("BENCH_CLOCK_MS", "20")) * 1e-3

    def _nvml_loop(self, nv, h):
        while self.running:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    why = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((float(sm), int(why)))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.running = True
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread:
            self.running = False
            self.thread.join(timeout=2)
            sm = [x[0] for x in self.samples]
            mask = 0
            for _, w in self.samples:
                mask |= w
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": sorted(n for n, b in self.BITS.items() if mask & b), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


class CpuArm:
    """The oracle port (oracle/ba_oracle.c, OpenMP) timed as the CPU arm: one LM iteration of the same fixed-work
    configuration per step (start linearisation + `pcg_iters` PCG iterations + trial evaluation + re-linearisation),
    CPU_THREADS pinned threads."""

    def __init__(self, P, implicit, dtype):
        from oracle.oracle import Oracle, lib
        lib().orc_set_num_threads(CPU_THREADS)
        self.O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype)
        self.x0 = P.x(dtype)
        self.implicit = implicit
        self.threads = self.O.num_threads()

    def step(self, pcg_iters):
        t0 = time.perf_counter()
        _, _, tr = self.O.lm_solve(self.x0, implicit=self.implicit, lm_max_iter=1, initial_region=LM_REGION, eps1=0.0,
                                   eps2=0.0, pcg_max_iter=pcg_iters, pcg_tol=0.0, pcg_refuse=1e300)
        return time.perf_counter() - t0, sum(t["pcg_iters"] for t in tr)

    def measure(self, pcg_iters, steps, warmup, budget_s):
        """Never fewer PCG iterations per step than asked; when the budget is short, FEWER STEPS are timed (and said
        so).  Returns None for value when not even one step fits."""
        t_first, n_first = self.step(pcg_iters)  # warm-up 1: also tells how many steps fit
        spent = t_first
        if t_first > budget_s:
            return dict(value=None, pcg=None, steps_timed=0, step_s=[t_first], spread=None, pcg_per_step=n_first,
                        fits=False, why=f"one step takes {t_first:.1f} s > budget {budget_s:.0f} s")
        for _ in range(max(warmup - 1, 0)):
            if spent + 2 * t_first > budget_s:
                break
            spent += self.step(pcg_iters)[0]
        times = []
        while len(times) < steps and (not times or spent + t_first <= budget_s):
            dt, n = self.step(pcg_iters)
            times.append(dt)
            spent += dt
        # repeat until two consecutive steps agree within 10 % (host noise), budget permitting
        while len(times) >= 2 and abs(times[-1] - times[-2]) > 0.1 * min(times[-2:]) and spent + t_first <= budget_s:
            dt, n = self.step(pcg_iters)
            times.append(dt)
            spent += dt
        t_step = statistics.median(times)
        # PCG rate: one more step with a fifth of the PCG iterations; the difference is pure PCG time
        pcg_rate = None
        n_small = max(pcg_iters // 5, 1)
        if n_small < pcg_iters and spent + t_first <= budget_s + 30:
            dt_small, got_small = self.step(n_small)
            if t_step > dt_small and n_first > got_small:
                pcg_rate = (n_first - got_small) / (t_step - dt_small)
        return dict(value=1.0 / t_step, pcg=pcg_rate, steps_timed=len(times), step_s=times,
                    spread=(max(times) - min(times)) / t_step, pcg_per_step=n_first, fits=len(times) >= steps, why=None)


def reference_gpu_solver(P, pcg_lo=10, pcg_hi=50, args_pcg_iters=50):
    """MegBA's own SchurPCGSolver::solve (src/solver/schur_pcg_solver.cu: cuBLAS matinvBatched, cuSPARSE CSR SpMV,
    cuBLAS dot/axpy, two host syncs per iteration) compiled for sm_100a from the reference sources via oracle/shim,
    timed on this GPU on the same problem with the same damped blocks.  Two fixed iteration counts give the time per
    PCG iteration by difference.  Returns None when oracle/_ref was not built (the reference is not on the GPU box)."""
    import ctypes as C
    lib = os.path.join(ROOT, "oracle", "_ref", "libmegba_ref.so")
    if not os.path.exists(lib):
        return None
    from oracle.oracle import Oracle
    L = C.CDLL(lib)
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, np.float64)
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J)
    HppD, _ = O.process_diag(Hpp, LM_REGION)
    HllD, _ = O.process_diag(Hll, LM_REGION)
    del r, J
    ptr = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    counts = [10, 20, 30, 40, 50]
    out = {}
    for rep in range(3):
        for n in counts:
            dx = np.zeros(9 * P.nC + 3 * P.nP)
            ms = C.c_double()
            rc = L.ref_solve_timed_f64(C.c_int64(P.nC), C.c_int64(P.nP), C.c_int64(P.nE), ptr(P.cam_idx), ptr(P.pt_idx),
                                       ptr(HppD), ptr(HllD), ptr(W), ptr(g), ptr(dx), C.c_int(n), C.c_double(0.0),
                                       C.c_double(1e300), C.byref(ms))
            if rc != 0:
                return None
            if rep > 0:  # the first sweep is warm-up (cuSPARSE buffers, clocks)
                out[n] = min(out.get(n, 1e30), ms.value)
    xs, ys = np.array(counts, float), np.array([out[n] for n in counts])
    A = np.vstack([xs, np.ones_like(xs)]).T
    (per_iter, intercept), res_, _, _ = np.linalg.lstsq(A, ys, rcond=None)
    rms = float(np.sqrt(np.mean((A @ np.array([per_iter, intercept]) - ys) ** 2)))
    res = {"pcg_iters_per_sec": 1e3 / per_iter, "ms_per_pcg_iter": float(per_iter), "fit_intercept_ms": float(intercept),
           "fit_rms_residual_ms": rms,
           "solve_ms": {str(k): v for k, v in out.items()},
           "what": "reference SchurPCGSolver::solve (schur_pcg_solver.cu, cuSPARSE/cuBLAS) built for sm_100a via "
                   "oracle/shim, same GPU and problem; least-squares slope of the solve time over 10..50 fixed "
                   "iterations (best of 2 per count after a warm-up sweep)"}
    del W, Hpp, Hll, HppD, HllD
    # the reference's whole LM loop (lm_algo.cu + its kernels + its solver), same fixed-work settings as our arm
    try:
        import re
        import tempfile
        n_lm = 4
        x = P.x(np.float64).copy()
        obs = np.ascontiguousarray(P.obs, np.float64)
        with tempfile.TemporaryDirectory() as td:
            log = os.path.join(td, "ref_lm.log")
            rc = L.ref_lm_solve_f64(C.c_int64(P.nC), C.c_int64(P.nP), C.c_int64(P.nE), ptr(P.cam_idx), ptr(P.pt_idx), ptr(obs),
                                    ptr(x), C.c_int(n_lm), C.c_double(LM_REGION), C.c_double(0.0), C.c_double(0.0),
                                    C.c_int(args_pcg_iters), C.c_double(0.0), C.c_double(1e300), C.c_char_p(log.encode()))
            text = open(log).read()
        el = [int(v) for v in re.findall(r"elapsed (\d+) ms", text)]
        if rc == 0 and len(el) >= n_lm + 1:
            steps_ms = [b - a for a, b in zip(el[:n_lm], el[1:n_lm + 1])]
            dt = max(min(steps_ms), 1) / 1e3  # its best iteration: the host-synchronous loop is sensitive to host noise
            res["lm_iters_per_sec"] = 1.0 / dt
            res["lm_ms_per_iter"] = steps_ms
            res["lm_what"] = (f"reference LMAlgo::solveCUDA (lm_algo.cu) + its kernels and solver via oracle/shim: {n_lm} LM "
                              f"iterations x {args_pcg_iters} PCG iterations, best iteration by the wall clock the reference prints itself")
            res["lm_trace"] = text.strip().splitlines()[: n_lm + 2]
    except Exception as exc:  # optional
        res["lm_error"] = repr(exc)
    return res


def run_reference(args, P, dtype):
    """CPU arm: the reference has no CPU execution path (every Device::CPU branch throws) and cannot be
    built here (Eigen/gflags absent), so this times the oracle port of its algorithm on the host's physical cores.
    Every timed step runs the full --pcg-iters; if K steps do not fit the time budget, fewer steps are timed."""
    arm = CpuArm(P, args.implicit, dtype)
    if args.cpu_arm_only > 0:
        m = arm.measure(args.pcg_iters, 2, 1, budget_s=args.cpu_arm_only)
        m["threads"] = arm.threads
        print(json.dumps({"cpu_arm": m}), flush=True)
        return
    m = arm.measure(args.pcg_iters, args.steps, args.warmup, budget_s=float(os.environ.get("BA_CPU_BUDGET_S", "200")))
    value = m["value"]
    sample = (f"{m['steps_timed']} x [1 LM iteration = 2 linearisations + {m['pcg_per_step']} PCG iterations + trial "
              f"evaluation] of the same {P.label} problem, median step; {arm.threads} pinned OpenMP threads "
              f"(OMP_PROC_BIND=close, OMP_PLACES=cores)" + ("" if m["fits"] else f"; {args.steps} steps asked: {m['why'] or 'time budget'}"))
    line = {
        "impl": "reference", "metric": "lm_iters_per_sec", "value": value, "unit": "LM it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": None if value is None else 1e3 / value,
        "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": data_label(P),
        "config": bench_config(args, P),
        "pcg_iters_per_sec": m["pcg"],
        "cpu_baseline": {"value": value, "unit": "LM it/s", "cores": arm.threads, "kind": "port", "sample": sample,
                         "pcg_iters_per_sec": m["pcg"], "steps_timed": m["steps_timed"], "step_seconds": m["step_s"],
                         "spread": m["spread"]},
        "e2e": {"value": value, "unit": "LM it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "consistency": {"fits_in_driver_run": bool(m["fits"]), "extrapolated": False, "threads": arm.threads,
                        "host_logical_cpus": os.cpu_count()},
        "gpu_launches": 0,
        "note": "MegBA has no CPU execution path (every Device::CPU branch throws); this arm is the oracle port of its "
                "algorithm on the host cores.  reference_gpu_solver below is MegBA's OWN CUDA code (oracle/shim build) on GPU 0.",
    }
    if dtype == np.float64 and not args.implicit and not args.no_cpu_baseline:
        try:
            import torch
            if torch.cuda.is_available():
                line["reference_gpu_solver"] = reference_gpu_solver(P, args_pcg_iters=args.pcg_iters)
        except Exception as exc:  # optional extra
            line["reference_gpu_solver"] = {"error": repr(exc)}
    print(json.dumps(line), flush=True)


def load_workload(name):
    """The real BAL file when $BAL_DIR holds it (problem-<nC>-<nP>-pre.txt); otherwise the
    synthetic problem of the same shape.  The JSON line's `data` / `config.workload` say which one ran."""
    from megba_b200.synthetic import BAL_SHAPES, make_named, read_bal
    bal_dir = os.environ.get("BAL_DIR")
    if bal_dir and name in BAL_SHAPES:
        nC, nP, _ = BAL_SHAPES[name]
        path = os.path.join(bal_dir, f"problem-{nC}-{nP}-pre.txt")
        if os.path.exists(path):
            return read_bal(path)
    return make_named(name)


def data_label(P):
    return "bal" if P.label.startswith("bal:") else "synthetic"


def bench_config(args, P):
    return {"workload": f"{P.label} ({P.nC} cameras, {P.nP} points, {P.nE} observations)",
            "compute_kind": "implicit" if args.implicit else "explicit",
            "layout": getattr(args, "layout", "auto"),
            "lm": {"tau": LM_REGION, "epsilon1": -1e-300, "epsilon2": -1, "note": "negative: both stopping tests disabled, exactly `steps` iterations"},
            "pcg": {"iters_per_lm_iter": args.pcg_iters, "tol": 0, "refuse_ratio": "inf"},
            "parallelism": f"point-sharded edges x{args.gpus}, cameras replicated",
            "l2": ("working set per PCG iteration (> 0.4 GB per rank at 1 GPU) exceeds the 126 MB L2; no explicit flush"
                   if P.nE * 70 / max(args.gpus, 1) > 126e6 else
                   "working set fits the 126 MB L2: the PCG pass is L2-resident by nature of the problem size")}


def run_ours(args, P, dtype):
    import torch
    import torch.distributed as dist

    from megba_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if ws != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={ws}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    if ws > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", local))

    layout = {"auto": capi.LAYOUT_AUTO, "ell": capi.LAYOUT_ELL, "tile": capi.LAYOUT_TILE}[args.layout]
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype, implicit=args.implicit, device=local,
                    rank=rank, world_size=ws, layout=layout)
    tile = args.layout == "tile"
    if ws > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.Solver.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        S.comm_init(bytes(uid.cpu().numpy().tobytes()))

    def barrier():
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if ws == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cam_in = torch.from_numpy(P.cam.astype(dtype)).pin_memory()
    pt_in = torch.from_numpy(P.pt.astype(dtype)).pin_memory()
    cam_out = torch.empty_like(cam_in).pin_memory()
    pt_out = torch.empty_like(pt_in).pin_memory()
    stream = torch.cuda.ExternalStream(S.stream())
    # negative epsilons: `||dx|| <= eps2 (||x|| + eps1)` and `||g||_inf <= eps1` can never fire, so exactly K iterations run
    kw = dict(initial_region=LM_REGION, eps1=-1e-300, eps2=-1.0, pcg_max_iter=args.pcg_iters, pcg_tol=0.0,
              pcg_refuse=1e300)

    # ---------------- resident arm: K iterations in one ba_lm_solve call
    S.set_params_ptr(cam_in.data_ptr(), pt_in.data_ptr())
    S.lm_solve(lm_max_iter=max(args.warmup, 3), **kw)  # warm-up (>= 3 iterations)
    S.set_params_ptr(cam_in.data_ptr(), pt_in.data_ptr())
    S.set_timing(8)  # event pairs around the hot kernels of every 8th PCG iteration (keeps event overhead out)
    S.reset_stats()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    trace = S.lm_solve(lm_max_iter=args.steps, **kw)
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    st = S.stats()
    S.set_timing(False)
    n_iter = len(trace) - 1
    assert n_iter == args.steps, (n_iter, args.steps)
    ms_total = max_over_ranks(dev_ms)
    wall = max_over_ranks(wall)
    lm_rate = args.steps / (ms_total * 1e-3)
    pcg_rate = st["pcg_iters"] / (max_over_ranks(st["pcg_loop_ms"]) * 1e-3)
    a_pcg, a_lin = S.algorithmic_bytes()
    peak, peak_src = peaks()
    k_ms = max_over_ranks(st["schur_apply_ms"] / max(st["schur_apply_timed"], 1))
    achieved = a_pcg / (k_ms * 1e-3) / 1e9
    traffic = None
    kname = (("k_schur_edge_full" if tile else "k_schur_ell_fact_full") + ("_f64" if dtype == np.float64 else "_f32"))
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and ws == 1 and args.workload == WORKLOAD and not args.implicit:
        with open(tpath) as fh:
            traffic = json.load(fh).get(kname)
    # SURVEY.md section 8d writes the per-observation figure for a kernel that streams the stored Hpl block
    # (27 s + 4 bytes); the default kernel streams 8 s + 4.  Both denominators are reported.
    its = np.dtype(dtype).itemsize
    nE_l, nP_l = P.nE / ws, P.nP / ws
    a_survey = nE_l * ((24 if args.implicit else 27) * its + 4) + nP_l * (9 * its + 4) + P.nC * 216 * its
    roofline = {"bound": "hbm",
                "kernel": ("k_schur_edge<FULL> (tile layout, W/J stream)" if tile else
                           "k_schur_ell<FACT,FULL> (ELL layout, factored stream: S*p edge pass of one PCG iteration)"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": a_pcg,
                "bytes_per_observation": {"stream": (a_pcg - P.nC * 216 * its) / max(nE_l, 1),
                                          "what": "this kernel's compulsory HBM bytes per observation incl. the per-point blocks (DESIGN.md section 3)"},
                "survey_8d_equivalent": {"algorithmic_bytes_per_launch": a_survey, "achieved": a_survey / (k_ms * 1e-3) / 1e9,
                                         "frac": a_survey / (k_ms * 1e-3) / 1e9 / peak,
                                         "what": "SURVEY.md section 8d bytes (a kernel streaming the stored 9x3 Hpl block) / this kernel's time"},
                "limiter": "LSU data pipe of the SM (l1tex__data_pipe_lsu_wavefronts ~73 % busy: gather of the camera vector, RED "
                           "scatter and its shared-memory transposition), not DRAM; see profiles/README.md",
                "avg_launch_ms": k_ms, "launches_timed": int(st["schur_apply_timed"]), "launches": st["schur_apply_calls"],
                "pcg_cam_kernel_avg_ms": st["pcg_cam_ms"] / max(st["pcg_cam_timed"], 1),
                "linearize": {"algorithmic_bytes": a_lin,
                              "avg_ms": st["linearize_ms"] / max(st["linearize_calls"], 1),
                              "achieved": a_lin / (st["linearize_ms"] / max(st["linearize_calls"], 1) * 1e-3) / 1e9}}
    # whole LM iteration against the same roofline: linearisation + trial evaluation +
    # (init pass + n_k PCG passes + back-substitution) x A_pcg, with the PCG iteration count actually run
    itemsize_ = np.dtype(dtype).itemsize
    a_eval = (P.nE * (8 + 2 * itemsize_) + P.nP * 6 * itemsize_ + P.nC * 18 * itemsize_) / ws
    n_lin = st["linearize_calls"]
    lm_bytes = n_lin * a_lin + args.steps * a_eval + (st["pcg_iters"] + 2 * args.steps) * a_pcg
    roofline["lm_iteration"] = {"algorithmic_bytes_per_step": lm_bytes / args.steps,
                                "achieved": lm_bytes / (ms_total * 1e-3) / 1e9, "unit": "GB/s",
                                "frac": lm_bytes / (ms_total * 1e-3) / 1e9 / peak,
                                "what": "all algorithmic bytes of the timed LM iterations / timed region (per rank)"}
    launches = st["launches"]

    # ---------------- end-to-end arm: K solve() calls on host buffers
    def e2e_step():
        S.set_params_ptr(cam_in.data_ptr(), pt_in.data_ptr())
        tr = S.lm_solve(lm_max_iter=1, **kw)
        S.get_params_ptr(cam_out.data_ptr(), pt_out.data_ptr())
        return tr[-1]["cost"]

    for _ in range(max(args.warmup, 3)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cost = e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None  # sampled through both timed regions (resident + end-to-end)
    itemsize = np.dtype(dtype).itemsize
    bytes_x = (9 * P.nC + 3 * P.nP) * itemsize
    e2e = {"value": args.steps / e2e_s, "unit": "LM it/s", "h2d_bytes_per_step": bytes_x,
           "d2h_bytes_per_step": bytes_x + 8, "ms_per_step": 1e3 * e2e_s / args.steps,
           "what": "set_params(H2D) + ba_lm_solve(max_iter=1) + get_params(D2H) per step", "final_cost": cost}

    # ---------------- README-parameter mode (README.md:58): stopping rules on, reported beside the fixed-work line
    readme = None
    if not args.no_readme_mode:
        rk = dict(initial_region=LM_REGION, eps1=1.0, eps2=1e-10, pcg_max_iter=100, pcg_tol=1e-1, pcg_refuse=1.0)
        S.set_params_ptr(cam_in.data_ptr(), pt_in.data_ptr())
        S.lm_solve(lm_max_iter=3, **rk)  # warm-up
        S.set_params_ptr(cam_in.data_ptr(), pt_in.data_ptr())
        S.reset_stats()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        tr = S.lm_solve(lm_max_iter=100, **rk)
        r1.record(stream)
        barrier()
        r_ms = max_over_ranks(r0.elapsed_time(r1))
        rst = S.stats()
        n_lm = len(tr) - 1
        readme = {"lm_iters": n_lm, "pcg_iters_total": rst["pcg_iters"], "ms_total": r_ms,
                  "lm_iters_per_sec": n_lm / (r_ms * 1e-3), "pcg_iters_per_sec_overall": rst["pcg_iters"] / (r_ms * 1e-3),
                  "accepted": sum(1 for t in tr[1:] if t["accepted"] == 1), "start_cost": tr[0]["cost"],
                  "final_cost": tr[-1]["cost"], "launches": rst["launches"],
                  "params": "max_iter 100, solver_max_iter 100, solver_tol 1e-1, solver_refuse_ratio 1, tau 1e4, epsilon1 1, "
                            "epsilon2 1e-10 (README.md:58); stops by the reference's own rules"}

    cpu = None
    if rank == 0 and ws == 1 and not args.no_cpu_baseline:
        # a child process, so that the OpenMP pinning of the CPU arm never touches this (GPU-driving) process
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--dtype", args.dtype,
               "--pcg-iters", str(args.pcg_iters), "--cpu-arm-only", os.environ.get("BA_CPU_BASELINE_BUDGET_S", "30")]
        if args.implicit:
            cmd.append("--implicit")
        out = subprocess.run(cmd, capture_output=True, text=True, env=pin_cpu_arm_env(dict(os.environ)))
        try:
            m = json.loads([l for l in out.stdout.splitlines() if l.startswith('{"cpu_arm"')][-1])["cpu_arm"]
            cpu = {"value": m["value"], "unit": "LM it/s", "cores": m["threads"], "kind": "port",
                   "pcg_iters_per_sec": m["pcg"], "steps_timed": m["steps_timed"], "step_seconds": m["step_s"],
                   "sample": f"median of {m['steps_timed']} x [1 LM iteration = 2 linearisations + {m['pcg_per_step']} PCG iterations] "
                             f"of the same problem after one warm-up step, oracle port (OpenMP, {m['threads']} pinned threads), in a child process"}
        except Exception as exc:
            cpu = {"value": None, "unit": "LM it/s", "cores": CPU_THREADS, "kind": "port", "error": repr(exc), "stderr": out.stderr[-300:]}

    ref_gpu = None
    if rank == 0 and ws == 1 and not args.no_cpu_baseline and dtype == np.float64 and not args.implicit:
        S.close()
        torch.cuda.empty_cache()
        try:
            ref_gpu = reference_gpu_solver(P, args_pcg_iters=args.pcg_iters)
        except Exception as exc:  # the checker library is optional
            ref_gpu = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": "lm_iters_per_sec", "value": lm_rate, "unit": "LM it/s", "n_gpus": ws, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": data_label(P),
            "config": bench_config(args, P),
            "pcg_iters_per_sec": pcg_rate, "pcg_iters_timed": st["pcg_iters"],
            "accepted_steps": sum(1 for t in trace[1:] if t["accepted"] == 1),
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "roofline": roofline, "readme_mode": readme, "cpu_baseline": cpu, "reference_gpu_solver": ref_gpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
            "final_cost": trace[-1]["cost"], "start_cost": trace[0]["cost"],
        }
        print(json.dumps(line), flush=True)
    S.close()
    if ws > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    dtype = np.float64 if args.dtype == "f64" else np.float32
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return  # rank 0 alone runs the CPU arm
    P = load_workload(args.workload)
    if args.impl == "reference":
        run_reference(args, P, dtype)
    else:
        run_ours(args, P, dtype)


if __name__ == "__main__":
    main()
