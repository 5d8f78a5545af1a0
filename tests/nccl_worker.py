"""Worker of tests/test_multi_nccl.py (one process per GPU, launched by torch.distributed.run): the in-library NCCL
exchange of libhadi.so against the single-GPU results, bit for bit.  Writes one JSON line per rank."""
import json
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
DIVS = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")     # plumbing only (id broadcast): the data path is the library's own NCCL
    hadi = ge.load_hadi()
    import importlib.util

    spec = importlib.util.spec_from_file_location("hadi_dist", os.path.join(ge.PKG, "hadi_dist.py"))
    hd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hd)
    ctx = hadi.Context(local)
    solo = hadi.Context(local)          # same device, no communicator: the single-GPU answer
    hd.attach_nccl(hadi, ctx, rank, world, dist=dist)
    assert ctx.comm_world == world
    mdl = hadi.make_model(**BASE)
    out = {"rank": rank}
    # prices: 333 American options with dividends, mixed step counts (uneven slices)
    K = [70.0 + 0.2 * k for k in range(333)]
    Ns = [10 + (k % 5) for k in range(333)]
    num = hadi.make_numerics(100, 50, 0.8, hadi.AMERICAN, hadi.CALL, hadi.DOUGLAS, DIVS)
    pts, n = hadi.make_points(K, 1.0, Ns)
    a = ctx.price_batch_sharded(mdl, num, pts, n)
    b = solo.price_batch(mdl, num, pts, n)["prices"]
    out["prices_equal"] = bool(np.array_equal(a, b))
    # Jacobian, all three modes
    nume = hadi.make_numerics(50, 25, 0.8)
    ptsj, nj = hadi.make_points([90.0 + k for k in range(21)], [1.0 + 0.25 * (k % 3) for k in range(21)],
                                [20 + 5 * (k % 3) for k in range(21)])
    for mode in (hadi.MODE_JACOBIAN, hadi.MODE_JACOBIAN_INTERP, hadi.MODE_JACOBIAN_CENTRAL):
        J, base = ctx.jacobian_batch_sharded(mdl, nume, ptsj, nj, mode)
        J1, base1 = solo.jacobian_batch_ex(mdl, nume, ptsj, nj, mode)
        out["jac_equal_%d" % mode] = bool(np.array_equal(J, J1) and np.array_equal(base, base1))
    # LM: the context's own communicator (comm=None) against the single-GPU run
    market = [hadi.bs_call(100.0, p.strike, 0.025, 0.2, p.maturity) for p in ptsj[:nj]]
    r = ctx.calibrate(mdl, nume, ptsj, nj, market, 6, 0.1 * math.sqrt(nj), 0.1 * (1 + math.log(nj)))
    r1 = solo.calibrate(mdl, nume, ptsj, nj, market, 6, 0.1 * math.sqrt(nj), 0.1 * (1 + math.log(nj)))
    out["lm_equal"] = bool(r["params"] == r1["params"] and r["final_error"] == r1["final_error"] and
                           r["iterations"] == r1["iterations"])
    # fewer items than ranks x 1: some ranks hold an empty slice
    pts1, n1 = hadi.make_points([100.0], 1.0, 8)
    out["single_item_equal"] = bool(np.array_equal(ctx.price_batch_sharded(mdl, nume, pts1, n1),
                                                   solo.price_batch(mdl, nume, pts1, n1)["prices"]))
    ctx.comm_finalize()
    print("NCCLWORKER " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
