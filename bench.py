#!/usr/bin/env python
"""bench.py — headline benchmark of the hadi engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch: BASELINE.json configs[1], 500 American options
with cash + proportional dividends, Douglas ADI + Ikonen-Toivanen, 101x51 grid, 50 time steps, per GPU
(weak scaling: every rank prices its own 500-option chain, no data-path collective).

Printed JSON (rank 0, one line):
  value            option solves / s over all GPUs, descriptors + grids already resident in HBM,
                   CUDA-event time of the kernel launches, max over ranks
  e2e              the same metric through the reference-facing C-ABI call hadi_price_batch() with HOST
                   buffers: descriptor build, H2D, kernel, D2H inside the timed region
  roofline         achieved algorithmic FP64 flop rate of the fused kernel vs the MEASURED un-fused
                   DMUL/DADD issue rate of this GPU (hadi_measure_fp64; MEASURED_PEAKS.json has no FP64
                   entry, its HBM figure is quoted for context)
  cpu_baseline     the reference's own code (oracle/_ref) on the host cores, bounded sample
  lm               wall-ms of the full LM calibration of BASELINE configs[2] (10x10 surface)
--impl reference times the reference's CPU implementation of the same path on all host cores.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Heston ADI option solves/sec (m1x m2 x N = 100x50x50, American + dividends)"
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
DIVS = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02, 0.02, 0.02, 0.02])
M1, M2, NSTEP, NOPT, THETA = 100, 50, 50, 500, 0.8
FLOPS_PER_POINT_STEP = 72          # algorithmic flops, American Douglas
P = (M1 + 1) * (M2 + 1)
WORKLOAD = "config2: 500 American calls with 4 cash+proportional dividends, Douglas ADI + Ikonen-Toivanen, 101x51 grid, N=50, strikes 70+0.12i"


def strikes_for(rank):
    # every rank prices its own chain (weak scaling); rank 0 is SURVEY C2 exactly
    return [70.0 + 0.12 * i + 0.01 * rank for i in range(NOPT)]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._stop_evt = threading.Event()

    def _run_nvml(self):
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
        while not self._stop_evt.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.rows.append([str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for _, b in bits])
            self._stop_evt.wait(0.02)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4), ("sw_power_cap", 5)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


_REF = None


class quiet_stdout:
    """The reference's host schemes print timings to stdout; bench.py must print exactly one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *a):
        os.dup2(self._saved, 1)
        os.close(self._null)
        os.close(self._saved)


def ref_lib():
    """oracle/_ref (the reference's own sources) with its league loop on ALL host cores.  torchrun exports
    OMP_NUM_THREADS=1 to its children, so the thread count is set through the library, not the environment,
    and the count reported is the one omp_get_max_threads() returns afterwards."""
    global _REF
    from oracle.reflib import RefLib, have_ref

    if _REF is None and have_ref(omp=True):
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)   # before libgomp is loaded
        R = RefLib(omp=True)
        _REF = (R, R.set_threads(os.cpu_count() or 1))
    return _REF


def cpu_baseline(sample_factor=None):
    """The reference's own sources (oracle/_ref, OpenMP over options outside the reference code) on the
    host cores, on a bounded sample of the same workload.  Timed: the reference entry point
    (compute_base_prices_american_dividends) alone — the driver's problem set-up (grids, U_0, the 12-array
    workspace) is outside the timed region, as it is outside the reference's own benchmark loops."""
    from oracle.reflib import OracleLib

    cores = os.cpu_count() or 1
    # default: the whole 500-option step (about 20 s of CPU work: 41 ms per solve and core)
    n = NOPT if sample_factor is None else max(8, min(NOPT, sample_factor * cores))
    strikes = strikes_for(0)[:n]
    ref = ref_lib()
    if ref is not None:
        R, cores = ref
        kind = "reference"

        def run():
            with quiet_stdout():
                p = R.solve_batch(strikes, NSTEP, 1.0 / NSTEP, m1=M1, m2=M2, theta=THETA, style=1, divs=DIVS, **BASE)["prices"]
            return p, R.last_compute_seconds()
    else:
        O = OracleLib()
        kind, cores = "port", 1
        n = 8
        strikes = strikes[:n]

        def run():
            t0 = time.perf_counter()
            p = O.price_batch(strikes, NSTEP, 1.0 / NSTEP, m1=M1, m2=M2, theta=THETA, style=1, divs=DIVS, **BASE)
            return p, time.perf_counter() - t0
    run()  # warm-up (page in, thread pool)
    t0 = time.perf_counter()
    prices, dt = run()
    wall = time.perf_counter() - t0
    return {"value": n / dt, "unit": "solves/s", "cores": cores, "kind": kind,
            "sample": "%d of the %d options of the step: %.2f s in the reference entry point on %d threads "
                      "(%.2f s wall with the driver's problem set-up)" % (n, NOPT, dt, cores, wall)}, prices, strikes


def bench_config():
    """`config` of the JSON line: identical keys and values on both arms."""
    return {"workload": WORKLOAD, "options_per_gpu": NOPT, "grid": "101x51", "time_steps": NSTEP,
            "l2": "256 MiB buffer written between timed iterations (L2 flush)",
            "timing": "CUDA events around each launch on the launching stream, summed over the steps, max over ranks"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    times = []
    base = None
    for it in range(args.warmup + args.steps):
        base, _, _ = cpu_baseline()
        if it >= args.warmup:
            times.append(base["value"])
    v = statistics.mean(times) if times else base["value"]
    base["value"] = v
    ms_per_step = NOPT / v * 1e3
    line = {"metric": METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": bench_config(),
            "reference_note": "each step = the whole 500-option workload on the host cores (reference entry point timed)",
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def lm_calibration(hadi, ctx, solo=None, world=1, rank=0, dist=None):
    """BASELINE configs[2]: LM calibration to a 10-strike x 10-maturity synthetic European surface."""
    K, T, N = c3_surface()
    market = [hadi.bs_call(100.0, k, 0.025, 0.2, t) for k, t in zip(K, T)]
    pts, n = hadi.make_points(K, T, N)
    out = {}
    for name, (m1, m2) in (("51x26", (50, 25)), ("101x51", (100, 50))):
        num = hadi.make_numerics(m1, m2, THETA)
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            res = ctx.calibrate(hadi.make_model(**BASE), num, pts, n, market, 15, 0.1 * math.sqrt(n),
                                0.1 * (1.0 + math.log(n)))   # sharded over the context's communicator when one is attached
            ms = (time.perf_counter() - t0) * 1e3
            best = ms if best is None else min(best, ms)
        gpu_ms = res["gpu_ms"]
        if dist is not None:
            import torch

            tb = torch.tensor([best, gpu_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)     # the slowest rank: its kernels hold the longest-maturity items
            best, gpu_ms = float(tb[0]), float(tb[1])
        out[name] = {"wall_ms": round(best, 3), "gpu_ms": round(gpu_ms, 3), "iterations": res["iterations"],
                     "pde_solves": res["pde_solves"], "converged": res["converged"],
                     "params": [repr(float(x)) for x in res["params"]], "final_error": repr(float(res["final_error"]))}
        if world > 1:
            # the same calibration on ONE GPU in the same run (rank 0; the others wait): strong-scaling efficiency
            if rank == 0:
                one = None
                for rep in range(3):
                    t0 = time.perf_counter()
                    r1 = solo.calibrate(hadi.make_model(**BASE), num, pts, n, market, 15, 0.1 * math.sqrt(n),
                                        0.1 * (1.0 + math.log(n)))
                    ms = (time.perf_counter() - t0) * 1e3
                    one = ms if one is None else min(one, ms)
                out[name]["single_gpu_wall_ms"] = round(one, 3)
                out[name]["strong_scaling_efficiency"] = round(one / (world * best), 4)
                out[name]["equals_single_gpu"] = bool(r1["params"] == res["params"] and r1["final_error"] == res["final_error"])
            dist.barrier()
        # opt-in solver-call schedule: the candidate is evaluated with its own Jacobian batch (one call per iteration,
        # 6n solves instead of 7n); same parameters and errors bit for bit, fewer and fuller launches
        sbest = None
        for rep in range(3):
            t0 = time.perf_counter()
            sres = ctx.calibrate(hadi.make_model(**BASE), num, pts, n, market, 15, 0.1 * math.sqrt(n),
                                 0.1 * (1.0 + math.log(n)), schedule=hadi.LM_SCHEDULE_SPECULATIVE)
            ms = (time.perf_counter() - t0) * 1e3
            sbest = ms if sbest is None else min(sbest, ms)
        sgpu = sres["gpu_ms"]
        if dist is not None:
            import torch

            tb = torch.tensor([sbest, sgpu], dtype=torch.float64, device="cuda")
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
            sbest, sgpu = float(tb[0]), float(tb[1])
        out[name]["speculative_schedule_optin"] = {
            "wall_ms": round(sbest, 3), "gpu_ms": round(sgpu, 3), "iterations": sres["iterations"],
            "pde_solves": sres["pde_solves"],
            "equals_reference_schedule": bool(sres["params"] == res["params"] and sres["final_error"] == res["final_error"]
                                              and sres["iterations"] == res["iterations"] and sres["lam"] == res["lam"])}
        try:   # golden trajectory of the reference itself (tests/golden/lm_more.json, oracle/make_golden.py more)
            G = json.load(open(os.path.join(ROOT, "tests", "golden", "lm_more.json")))["config3_" + name]
            out[name]["equals_reference"] = bool(out[name]["params"] == G["params"] and
                                                 out[name]["final_error"] == G["final_error"] and
                                                 res["iterations"] == G["iterations"])
        except Exception:
            out[name]["equals_reference"] = None
    # opt-in: V0 column of the Jacobian interpolated on the base solve (5 solves per point; not the reference's
    # trajectory — reported beside the parity run, never instead of it)
    num = hadi.make_numerics(50, 25, THETA)
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        res = ctx.calibrate(hadi.make_model(**BASE), num, pts, n, market, 15, 0.1 * math.sqrt(n),
                            0.1 * (1.0 + math.log(n)), jac_mode=hadi.MODE_JACOBIAN_INTERP)
        ms = (time.perf_counter() - t0) * 1e3
        best = ms if best is None else min(best, ms)
    out["51x26_interpolated_v0_optin"] = {"wall_ms": round(best, 3), "gpu_ms": round(res["gpu_ms"], 3),
                                          "iterations": res["iterations"], "pde_solves": res["pde_solves"],
                                          "converged": res["converged"]}
    return out


def c5_chain():
    """SURVEY C5 exactly: K_i = 50 + 0.01 i, maturities cycling {0.25, 0.5, 1, 2}, N = max(20, int(20 T)),
    European Douglas on the 101x51 grid."""
    cyc = (0.25, 0.5, 1.0, 2.0)
    K = [50.0 + 0.01 * i for i in range(10000)]
    T = [cyc[i % 4] for i in range(10000)]
    N = [max(20, int(20 * t)) for t in T]
    return K, T, N


def c3_surface():
    """BASELINE configs[2] / SURVEY C3: 10 strikes x 10 maturities, market = Black-Scholes at 20 % vol."""
    mats = [1.0 + i * 0.25 if i < 8 else 3.0 + (i - 8) * 0.5 for i in range(10)]
    K, T, N = [], [], []
    for Tm in mats:
        for s in range(10):
            K.append(95.0 + 1.0 * s)
            T.append(Tm)
            N.append(max(20, int(Tm * 20)))
    return K, T, N


def _digest(a):
    import hashlib
    import numpy as np

    return hashlib.sha256((np.ascontiguousarray(a, dtype=np.float64) + 0.0).tobytes()).hexdigest()[:16]


def sharded_chains(hadi, ctx, solo, torch, dev, rank, world, dist):
    """Strong scaling, one workload split over all ranks through the library's own exchange (hadi_*_sharded: items
    block-partitioned by cost, every rank's kernel publishes into the gather buffer, one ncclAllGather on the context's
    stream, one D2H): (a) the 500 American+dividend options of config 2 (the "500 options in <= 2 ms on 8 GPUs"
    target), (b) the SURVEY C5 chain (10 000 European options, mixed maturities), (c) the finite-difference Jacobian of
    the C3 surface (100 points x 6 solves, 101x51).  Wall time of the one-call entry point (host buffers in, host
    buffers out, descriptor build and H2D included), from a barrier to the results on the host, max over ranks, best of
    5; `digest` is a sha256 prefix of the results — it must not change with the number of ranks."""
    import numpy as np

    mdl = hadi.make_model(**BASE)
    out = {}
    K5, T5, N5 = c5_chain()
    K3, T3, N3 = c3_surface()
    cases = (("config2_500_sharded", hadi.make_numerics(M1, M2, THETA, hadi.AMERICAN, hadi.CALL, hadi.DOUGLAS, DIVS),
              hadi.make_points([70.0 + 0.12 * i for i in range(NOPT)], 1.0, NSTEP), hadi.MODE_PRICE),
             ("c5_chain_10k_sharded", hadi.make_numerics(M1, M2, THETA), hadi.make_points(K5, T5, N5), hadi.MODE_PRICE),
             ("c3_jacobian_600_sharded", hadi.make_numerics(M1, M2, THETA), hadi.make_points(K3, T3, N3), hadi.MODE_JACOBIAN))

    def run(c, num, pts, n, mode, sharded):
        if mode == hadi.MODE_PRICE:
            return c.price_batch_sharded(mdl, num, pts, n) if sharded else c.price_batch(mdl, num, pts, n)["prices"]
        J, b = c.jacobian_batch_sharded(mdl, num, pts, n) if sharded else c.jacobian_batch(mdl, num, pts, n)
        return np.concatenate([J.ravel(), b])

    for name, num, (pts, n), mode in cases:
        items = n * hadi.ITEMS_PER_OPTION[mode]
        best = None
        for rep in range(7):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            vals = run(ctx, num, pts, n, mode, world > 1)
            ms = (time.perf_counter() - t0) * 1e3
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep >= 2:
                best = float(t[0]) if best is None else min(best, float(t[0]))
        out[name] = {"items": int(items), "wall_ms": round(best, 3), "solves_per_s": round(items / (best * 1e-3), 1),
                     "checksum": float(np.sum(vals)), "digest": _digest(vals)}
        if world > 1:
            # the same workload on ONE GPU in the same run (rank 0, the others wait): strong-scaling efficiency
            if rank == 0:
                one = None
                for rep in range(5):
                    t0 = time.perf_counter()
                    v1 = run(solo, num, pts, n, mode, False)
                    ms = (time.perf_counter() - t0) * 1e3
                    if rep >= 2:
                        one = ms if one is None else min(one, ms)
                out[name]["single_gpu_ms"] = round(one, 3)
                out[name]["strong_scaling_efficiency"] = round(one / (world * best), 4)
                out[name]["equals_single_gpu"] = bool(np.array_equal(v1, vals))
            dist.barrier()
    return out


def c5_cpu_baseline():
    """The reference's own code on the host cores on a bounded sample of the C5 chain (every 50th option:
    200 options covering all four maturities; compute_base_prices_multi_maturity timed alone)."""
    import numpy as np

    ref = ref_lib()
    if ref is None:
        return None
    R, cores = ref
    K, T, N = c5_chain()
    idx = list(range(0, 10000, 50))
    Ks, Ts = np.array([K[i] for i in idx]), np.array([T[i] for i in idx])
    Ns = np.array([N[i] for i in idx], dtype=np.int32)
    with quiet_stdout():
        R.solve_batch(Ks[:cores], Ns[:cores], Ts[:cores] / Ns[:cores], maturities=Ts[:cores], m1=M1, m2=M2, theta=THETA, multi=1, **BASE)
        p = R.solve_batch(Ks, Ns, Ts / Ns, maturities=Ts, m1=M1, m2=M2, theta=THETA, multi=1, **BASE)["prices"]
    dt = R.last_compute_seconds()
    return {"value": len(idx) / dt, "unit": "solves/s", "cores": cores, "kind": "reference",
            "sample": "every 50th option of the chain (200 options), %.2f s in the reference entry point" % dt}, p, idx


def config4_block(hadi, ctx, peaks, with_cpu):
    """BASELINE configs[3]: European Craig-Sneyd on the 401x201 grid, N = 200 — beyond shared memory, so U, Y and
    the Craig-Sneyd stage arrays live in global scratch.  Batches of up to 48 such solves run on the wide kernel (one
    solve on a team of co-resident CTAs, line solves out of shared memory), larger ones one CTA per solve.  Two bounds
    are reported: the HBM roofline of SURVEY 8(d) (algorithmic bytes = 9 arrays x 8 B x P per step over the kernel
    time, against the measured copy bandwidth), which is what bounds MANY solves, and for the few-solve regime the
    dependent-chain floor of the line solves (per step two A1 sweeps of m1 nodes x 7 dependent FP64 operations and two
    A2 sweeps of m2+1 nodes x 7, at the measured dependent-issue interval of `dep_cycles` cycles), which no amount of
    parallel hardware shortens while the sweeps stay sequential (bit parity with the reference's Thomas order)."""
    m1, m2, N = 400, 200, 200
    Pl = (m1 + 1) * (m2 + 1)
    mdl = hadi.make_model(**BASE)
    num = hadi.make_numerics(m1, m2, THETA, hadi.EUROPEAN, hadi.CALL, hadi.CRAIG_SNEYD, None)
    out = {"workload": "config4: European call, Craig-Sneyd, 401x201 grid, N=200, strikes 100+0.1k"}
    hbm = peaks.get("hbm_gbs") or 6552.3
    dep_cycles, sm_ghz = 8.1, 1.965   # tools/ubench_lat.cu on B200; SM clock under load
    chain_floor_ms = N * 2 * (m1 * 7 + (m2 + 1) * 7) * dep_cycles / (sm_ghz * 1e6)
    for nopt in (1, 8, 37, 148):
        pts, n = hadi.make_points([100.0 + 0.1 * k for k in range(nopt)], 1.0, N)
        bt = ctx.batch(mdl, num, pts, n)
        ts = []
        for r in range(3 if nopt < 148 else 2):
            bt.launch()
            v = bt.fetch()
            ts.append(bt.elapsed_ms())
        ms = min(ts)
        byts = nopt * N * Pl * 8 * 9
        variant, ctas, per_solve = bt.kernel_info
        ent = {"solves": nopt, "ms": round(ms, 3), "ms_per_solve": round(ms / nopt, 3),
               "solves_per_s": round(nopt / (ms * 1e-3), 2),
               "kernel": {"variant": variant, "ctas": ctas, "ctas_per_solve": per_solve},
               "roofline": {"bound": "hbm", "achieved": round(byts / ms / 1e6, 1), "peak": hbm,
                            "unit": "GB/s", "frac": round(byts / ms / 1e6 / hbm, 4)},
               "price0": repr(float(v[0]))}
        if nopt == 1:
            ent["latency_bound"] = {"bound": "dependent FP64 chain of the line solves", "floor_ms": round(chain_floor_ms, 3),
                                    "frac": round(chain_floor_ms / ms, 4)}
        out["n%d" % nopt] = ent
        bt.destroy()
    out["golden_price_K100"] = "8.8920027296371611"
    if with_cpu:
        ref = ref_lib()
        if ref is not None:
            R = ref[0]
            t0 = time.perf_counter()
            with quiet_stdout():
                pr = R.host_scheme(1, K=100.0, T=1.0, m1=m1, m2=m2, N=N, theta=THETA, **BASE)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "solves/s", "cores": 1, "kind": "reference",
                                   "sample": "1 solve through the reference's CS_scheme_shuffled (host matrix classes, serial), %.2f s" % dt,
                                   "price": repr(float(pr)), "parity": repr(float(pr)) == out["n1"]["price0"]}
    return out


def instance_sweep(hadi, ctx):
    """The reference's own benchmark protocol (src/perfomance_test.cpp:46-57): European calls, strike 85, 51x26 grid,
    N = 20, instance counts {1, 10, 20, 50, 100, 200, 300, 500}, mean of 10 runs — end to end through
    hadi_price_batch with host buffers (the latency regime of small batches)."""
    mdl = hadi.make_model(**BASE)
    num = hadi.make_numerics(50, 25, THETA)
    out = {}
    for ninst in (1, 10, 20, 50, 100, 200, 300, 500):
        pts, n = hadi.make_points([85.0] * ninst, 1.0, 20)
        for _ in range(3):
            ctx.price_batch(mdl, num, pts, n)
        t0 = time.perf_counter()
        for _ in range(10):
            r = ctx.price_batch(mdl, num, pts, n)
        ms = (time.perf_counter() - t0) * 1e3 / 10
        out[str(ninst)] = {"ms": round(ms, 4), "ms_per_instance": round(ms / ninst, 5),
                           "solves_per_s": round(ninst / (ms * 1e-3), 1)}
    out["price_K85"] = repr(float(r["prices"][0]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline + LM + sharded only (no config 4, no instance sweep)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    hadi = ge.load_hadi()
    if not os.path.exists(hadi.LIB_PATH):
        raise SystemExit("libhadi.so is missing: run __graft_entry__.build() (there is no fallback path)")
    dev = torch.device("cuda", local_rank)
    ctx = hadi.Context(local_rank)
    mdl = hadi.make_model(**BASE)
    num = hadi.make_numerics(M1, M2, THETA, hadi.AMERICAN, hadi.CALL, hadi.DOUGLAS, DIVS)
    strikes = strikes_for(rank)
    pts, n = hadi.make_points(strikes, 1.0, NSTEP)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2

    def l2_flush():
        flush.zero_()
        torch.cuda.synchronize(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput (value) -----------------------------------------------------
    bt = ctx.batch(mdl, num, pts, n)
    for _ in range(args.warmup):
        bt.launch()
        bt.fetch()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.kernel_launches
    barrier()
    kernel_ms = []
    for _ in range(args.steps):
        l2_flush()
        bt.launch()
        prices = bt.fetch()
        kernel_ms.append(bt.elapsed_ms())      # CUDA events on the launching stream
    barrier()
    launches = ctx.kernel_launches - launches0
    total_ms = float(sum(kernel_ms))

    # ---- end to end through the C-ABI call with host buffers (e2e) ---------------------------------
    for _ in range(2):
        ctx.price_batch(mdl, num, pts, n)
    h0, d0 = ctx.transfer_bytes()
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        l2_flush()
        t0 = time.perf_counter()
        out = ctx.price_batch(mdl, num, pts, n)   # descriptors + H2D + kernel + D2H + sync
        e2e_s += time.perf_counter() - t0
    barrier()
    h1, d1 = ctx.transfer_bytes()
    launches += args.steps
    clocks = sampler.stop() if sampler else None

    tt = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_ms_max = float(tt[0]), float(tt[1])
    value = world * NOPT * args.steps / (total_ms_max * 1e-3)
    e2e_value = world * NOPT * args.steps / (e2e_ms_max * 1e-3)

    # ---- LM calibration (configs[2]) and strong-scaling workloads; sharded over the ranks when world > 1 through the
    # library's own NCCL communicator (hadi_comm_init): no Python in the exchange
    solo = None
    if dist is not None:
        import importlib.util

        spec = importlib.util.spec_from_file_location("hadi_dist", os.path.join(ge.PKG, "hadi_dist.py"))
        hd = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(hd)
        hd.attach_nccl(hadi, ctx, rank, world, dist=dist, device=dev)
        solo = hadi.Context(local_rank)     # no communicator: the single-GPU reference of the same run
    lm = lm_calibration(hadi, ctx, solo, world, rank, dist)
    sharded = sharded_chains(hadi, ctx, solo, torch, dev, rank, world, dist)
    if dist is not None:
        dist.barrier()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    config4 = sweep = None
    if rank == 0 and world == 1 and not args.quick:
        config4 = config4_block(hadi, ctx, peaks, with_cpu=not args.no_cpu_baseline)
        sweep = instance_sweep(hadi, ctx)

    if rank == 0:
        unfused, fma, dep_ns = hadi.measure_fp64(local_rank)
        ms_kernel = total_ms / args.steps
        flops = NOPT * NSTEP * P * FLOPS_PER_POINT_STEP
        achieved = flops / (ms_kernel * 1e-3) / 1e12
        traffic = None
        try:   # DRAM bytes of one launch of this kernel from the latest committed ncu --set full capture
            import glob

            latest = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))[-1]
            traffic = json.load(open(latest))["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "fp64", "achieved": achieved, "peak": unfused, "unit": "TFLOP/s",
                    "frac": achieved / unfused,
                    "traffic": traffic,
                    "note": "FP64-pipe bound, SMEM-resident (not hbm/tensor): peak = un-fused DMUL+DADD issue rate "
                            "measured live by hadi_measure_fp64 (parity forbids FMA; DFMA rate %.1f TFLOP/s); "
                            "algorithmic flops = %d options x %d steps x %d nodes x %d; algorithmic HBM bytes per "
                            "launch = %d (descriptors+grids in, prices out); MEASURED_PEAKS hbm_gbs=%s for context"
                            % (fma, NOPT, NSTEP, P, FLOPS_PER_POINT_STEP, (h1 - h0) // args.steps + (d1 - d0) // args.steps,
                               peaks.get("hbm_gbs"))}
        line = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": bench_config(),
                "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": (h1 - h0) // args.steps,
                        "d2h_bytes_per_step": (d1 - d0) // args.steps, "ms_per_step": e2e_ms_max / args.steps,
                        "api": "hadi_price_batch (C ABI, host buffers)"},
                "gpu_launches": int(launches),
                "roofline": roofline,
                "clocks": clocks,
                "lm": lm,
                "sharded": sharded,
                "config4": config4,
                "instance_sweep_51x26x20": sweep,
                "grid_point_steps_per_s": value * NSTEP * P}
        if world == 1 and not args.no_cpu_baseline:
            base, ref_prices, ref_strikes = cpu_baseline()
            line["cpu_baseline"] = base
            import numpy as np

            line["cpu_baseline"]["parity_on_sample"] = bool(np.array_equal(np.asarray(ref_prices), np.asarray(out["prices"][:len(ref_strikes)])))
            c5 = c5_cpu_baseline()
            if c5 is not None:
                # parity of the sample against the GPU chain, and the CPU rate beside the sharded GPU rate
                K5, T5, N5 = c5_chain()
                idx = c5[2]
                pts5, n5 = hadi.make_points([K5[i] for i in idx], [T5[i] for i in idx], [N5[i] for i in idx])
                g5 = ctx.price_batch(mdl, hadi.make_numerics(M1, M2, THETA), pts5, n5)["prices"]
                c5[0]["parity_on_sample"] = bool(np.array_equal(np.asarray(c5[1]), np.asarray(g5)))
                line["sharded"]["c5_chain_10k_sharded"]["cpu_baseline"] = c5[0]
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
