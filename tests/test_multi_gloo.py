"""N > 1 host logic on CPU: two gloo ranks shard a batch with the C partitioner, solve their slices
(with the oracle standing in for the per-rank GPU solve), all-gather, and must reproduce the
single-process result bit for bit on every rank."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import BASE, ROOT


def _worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge
    from oracle.reflib import OracleLib

    hadi = ge.load_hadi()
    import importlib.util

    spec = importlib.util.spec_from_file_location("hadi_dist", os.path.join(ge.PKG, "hadi_dist.py"))
    hd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hd)
    O = OracleLib()
    strikes = [90.0, 95.0, 100.0, 105.0, 110.0]
    mats = [1.0, 1.0, 2.0, 0.5, 1.5]
    Ns = [8, 8, 16, 6, 12]
    num = hadi.make_numerics(20, 10, 0.8)
    pts, n = hadi.make_points(strikes, mats, Ns)
    eps = 1e-6

    def local_solve(b, e):
        # item = option*6 + column; column 0 base, 1..4 kappa/eta/sigma/rho + eps, 5 the V0 + eps grid
        vals = []
        for item in range(b, e):
            k, col = divmod(item, 6)
            m = dict(BASE)
            if col == 1: m["kappa"] += eps
            if col == 2: m["eta"] += eps
            if col == 3: m["sigma"] += eps
            if col == 4: m["rho"] += eps
            if col == 5: m["V0"] = BASE["V0"] + eps
            vals.append(O.solve(strikes[k], Ns[k], mats[k] / Ns[k], m1=20, m2=10, theta=0.8, want_U=False,
                                want_lambda=False, **m)["price"])
        return np.array(vals)

    vals = hd.solve_items_sharded(hadi, num, pts, n, hadi.MODE_JACOBIAN, local_solve, rank, world, dist=dist)
    J, base = hadi.jacobian_assemble(vals, eps)
    delta = hadi.lm_update(J, np.linspace(-0.1, 0.1, n), 0.01)
    # opt-in Jacobian with the V0 column interpolated on the base solve: 5 items per option, THREE values per
    # item (price, U at the two v-rows bracketing V0 + eps on the S0 column) — the gather counts scale with it
    lo, hi, wgt = hadi.jacobian_v0_weight(10, BASE["V0"], eps)

    def local_solve_interp(b, e):
        vals = []
        for item in range(b, e):
            k, col = divmod(item, 5)
            m = dict(BASE)
            if col == 1: m["kappa"] += eps
            if col == 2: m["eta"] += eps
            if col == 3: m["sigma"] += eps
            if col == 4: m["rho"] += eps
            o = O.solve(strikes[k], Ns[k], mats[k] / Ns[k], m1=20, m2=10, theta=0.8, want_U=True, want_lambda=False, **m)
            sg, _ = hadi.grid(20, 10, strikes[k], BASE["S0"], BASE["V0"])
            i_s = int(np.argmax(np.abs(sg - BASE["S0"]) < 1e-10))
            U = np.asarray(o["U"]).reshape(11, 21)
            vals += [o["price"], U[lo, i_s], U[hi, i_s]]
        return np.array(vals)

    vi = hd.solve_items_sharded(hadi, num, pts, n, hadi.MODE_JACOBIAN_INTERP, local_solve_interp, rank, world, dist=dist)
    Ji, bi = hadi.jacobian_assemble_ex(vi, hadi.MODE_JACOBIAN_INTERP, eps, wgt)
    full = local_solve_interp(0, 5 * n)
    assert np.array_equal(vi, full) and np.array_equal(bi, base) and np.array_equal(Ji[:, :4], J[:, :4])
    for k in range(n):
        o = full[15 * k:15 * k + 3]
        assert Ji[k, 4] == ((o[1] + wgt * (o[2] - o[1])) - o[0]) / eps
    q.put((rank, vals.tolist(), J.tolist(), base.tolist(), delta.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_jacobian_matches_single_process(hadi, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1:] == res[1][1:]  # ranks are bit-identical without a broadcast
    strikes = [90.0, 95.0, 100.0, 105.0, 110.0]
    mats = np.array([1.0, 1.0, 2.0, 0.5, 1.5])
    Ns = np.array([8, 8, 16, 6, 12], dtype=np.int32)
    Jo, bo = oracle.jacobian_batch(strikes, Ns, mats / Ns, m1=20, m2=10, theta=0.8, **BASE)
    assert np.array_equal(np.array(res[0][2]), Jo)
    assert np.array_equal(np.array(res[0][3]), bo)
