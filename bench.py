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


def cpu_baseline(sample_factor=None):
    """The reference's own sources (oracle/_ref, OpenMP over options outside the reference code) on the
    host cores, on a bounded sample of the same workload."""
    from oracle.reflib import OracleLib, RefLib, have_ref

    cores = os.cpu_count() or 1
    # default: the whole 500-option step (about 20 s of CPU work: 41 ms per solve and core)
    n = NOPT if sample_factor is None else max(8, min(NOPT, sample_factor * cores))
    strikes = strikes_for(0)[:n]
    if have_ref(omp=True):
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        R = RefLib(omp=True)
        kind = "reference"

        def run():
            return R.solve_batch(strikes, NSTEP, 1.0 / NSTEP, m1=M1, m2=M2, theta=THETA, style=1, divs=DIVS, **BASE)["prices"]
    else:
        O = OracleLib()
        kind, cores = "port", 1
        n = 8
        strikes = strikes[:n]

        def run():
            return O.price_batch(strikes, NSTEP, 1.0 / NSTEP, m1=M1, m2=M2, theta=THETA, style=1, divs=DIVS, **BASE)
    run()  # warm-up (page in, thread pool)
    t0 = time.perf_counter()
    prices = run()
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "solves/s", "cores": cores, "kind": kind,
            "sample": "%d of the %d options of the step, %.2f s wall" % (n, NOPT, dt)}, prices, strikes


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
    line = {"metric": METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "note": "each step = bounded sample of the workload on the host cores"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def lm_calibration(hadi, ctx, comm=None):
    """BASELINE configs[2]: LM calibration to a 10-strike x 10-maturity synthetic European surface."""
    mats = [1.0 + i * 0.25 if i < 8 else 3.0 + (i - 8) * 0.5 for i in range(10)]
    K, T, N = [], [], []
    for Tm in mats:
        for s in range(10):
            K.append(95.0 + 1.0 * s)
            T.append(Tm)
            N.append(max(20, int(Tm * 20)))
    market = [hadi.bs_call(100.0, k, 0.025, 0.2, t) for k, t in zip(K, T)]
    pts, n = hadi.make_points(K, T, N)
    out = {}
    for name, (m1, m2) in (("51x26", (50, 25)), ("101x51", (100, 50))):
        num = hadi.make_numerics(m1, m2, THETA)
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            res = ctx.calibrate(hadi.make_model(**BASE), num, pts, n, market, 15, 0.1 * math.sqrt(n),
                                0.1 * (1.0 + math.log(n)), comm=comm)
            ms = (time.perf_counter() - t0) * 1e3
            best = ms if best is None else min(best, ms)
        out[name] = {"wall_ms": round(best, 3), "gpu_ms": round(res["gpu_ms"], 3), "iterations": res["iterations"],
                     "pde_solves": res["pde_solves"], "converged": res["converged"]}
    # opt-in: V0 column of the Jacobian interpolated on the base solve (5 solves per point; not the reference's
    # trajectory — reported beside the parity run, never instead of it)
    num = hadi.make_numerics(50, 25, THETA)
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        res = ctx.calibrate(hadi.make_model(**BASE), num, pts, n, market, 15, 0.1 * math.sqrt(n),
                            0.1 * (1.0 + math.log(n)), comm=comm, jac_mode=hadi.MODE_JACOBIAN_INTERP)
        ms = (time.perf_counter() - t0) * 1e3
        best = ms if best is None else min(best, ms)
    out["51x26_interpolated_v0_optin"] = {"wall_ms": round(best, 3), "gpu_ms": round(res["gpu_ms"], 3),
                                          "iterations": res["iterations"], "pde_solves": res["pde_solves"],
                                          "converged": res["converged"]}
    return out


def sharded_chains(hadi, ctx, torch, dev, rank, world, dist):
    """BASELINE target and configs[4]: (a) the 500 American+dividend options of config 2 priced ONCE, sharded over
    all ranks (strong scaling: the "500 options in <= 2 ms on 8 GPUs" target), and (b) a 10 000-option European
    chain (101x51, N=50) sharded the same way.  Items are block-partitioned by cost (hadi_partition); every rank
    solves its slice and the values are all-gathered (NCCL) — wall time from the barrier before the launches to
    the gathered values on the host, max over ranks, best of 5."""
    import importlib.util
    import numpy as np
    import __graft_entry__ as ge

    spec = importlib.util.spec_from_file_location("hadi_dist", os.path.join(ge.PKG, "hadi_dist.py"))
    hd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hd)
    mdl = hadi.make_model(**BASE)
    out = {}
    cases = (("config2_500_sharded", hadi.make_numerics(M1, M2, THETA, hadi.AMERICAN, hadi.CALL, hadi.DOUGLAS, DIVS),
              [70.0 + 0.12 * i for i in range(NOPT)]),
             ("chain_10k_sharded", hadi.make_numerics(M1, M2, THETA), [60.0 + 0.008 * i for i in range(10000)]))
    for name, num, strikes in cases:
        pts, n = hadi.make_points(strikes, 1.0, NSTEP)
        costs = hadi.item_costs(num, pts, n, hadi.MODE_PRICE)
        sl = hd.slices(hadi, costs, world)
        b, e = sl[rank]
        bt = ctx.batch(mdl, num, pts, n, begin=b, end=e) if e > b else None
        counts = [x[1] - x[0] for x in sl]
        best = None
        for rep in range(7):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            mine = np.zeros(0)
            if bt is not None:
                bt.launch()
                mine = bt.fetch()
            vals = mine if world == 1 else hd.allgather_values(mine, counts, dist=dist, device=dev)
            ms = (time.perf_counter() - t0) * 1e3
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep >= 2:
                best = float(t[0]) if best is None else min(best, float(t[0]))
        out[name] = {"options": len(strikes), "wall_ms": round(best, 3), "solves_per_s": round(len(strikes) / (best * 1e-3), 1),
                     "checksum": float(np.sum(vals))}
        if bt is not None:
            bt.destroy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    # ---- LM calibration (configs[2]); sharded over the ranks when world > 1 -------------------------
    comm = None
    if dist is not None:
        import importlib.util

        spec = importlib.util.spec_from_file_location("hadi_dist", os.path.join(ge.PKG, "hadi_dist.py"))
        hd = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(hd)
        comm = hd.make_comm(hadi, rank, world, device=dev, dist=dist)
    lm = lm_calibration(hadi, ctx, comm)
    sharded = sharded_chains(hadi, ctx, torch, dev, rank, world, dist)
    if dist is not None:
        dist.barrier()

    if rank == 0:
        unfused, fma, dep_ns = hadi.measure_fp64(local_rank)
        ms_kernel = total_ms / args.steps
        flops = NOPT * NSTEP * P * FLOPS_PER_POINT_STEP
        achieved = flops / (ms_kernel * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        try:   # DRAM bytes of one launch of this kernel from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["dram_bytes_per_launch"]
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
                "config": {"workload": WORKLOAD, "options_per_gpu": NOPT, "grid": "101x51", "time_steps": NSTEP,
                           "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                           "timing": "CUDA events around each launch on the launching stream, summed over the steps, max over ranks"},
                "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": (h1 - h0) // args.steps,
                        "d2h_bytes_per_step": (d1 - d0) // args.steps, "ms_per_step": e2e_ms_max / args.steps,
                        "api": "hadi_price_batch (C ABI, host buffers)"},
                "gpu_launches": int(launches),
                "roofline": roofline,
                "clocks": clocks,
                "lm": lm,
                "sharded": sharded,
                "grid_point_steps_per_s": value * NSTEP * P}
        if world == 1 and not args.no_cpu_baseline:
            base, ref_prices, ref_strikes = cpu_baseline()
            line["cpu_baseline"] = base
            import numpy as np

            line["cpu_baseline"]["parity_on_sample"] = bool(np.array_equal(np.asarray(ref_prices), np.asarray(out["prices"][:len(ref_strikes)])))
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
