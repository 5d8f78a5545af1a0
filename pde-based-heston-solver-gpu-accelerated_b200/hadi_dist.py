"""Multi-GPU plumbing: one process per GPU, torch.distributed for the (tiny) exchange.

The path shards embarrassingly: work items — options, or option x Jacobian column —
are independent PDE solves.  Each rank solves a contiguous, cost-balanced slice (hadi_partition) on its
own GPU; the only exchange is one all-gather of the item values (8 bytes per item).  Every rank then
assembles J and runs the replicated LM update in the same summation order, so ranks stay bit-identical
without a broadcast.  There is no data-path collective inside the solver.
"""
import ctypes as C

import numpy as np


def slices(hadi, costs, world):
    """[(begin, end)] per rank from the C partitioner."""
    return [hadi.partition(costs, world, r) for r in range(world)]


_bufs = {}


def allgather_values(mine, counts, dist=None, device=None):
    """All-gather variable-length float64 slices in rank order: one padded all_gather_into_tensor and one
    device-to-host copy per call; the staging tensors are cached per (longest slice, world, device)."""
    import torch

    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size()
    if not sum(counts):
        return np.zeros(0)
    mx = max(max(counts), 1)
    key = (mx, world, str(device))
    if key not in _bufs:
        pin = device is not None and torch.device(device).type == "cuda"
        _bufs[key] = (torch.zeros(mx, dtype=torch.float64, pin_memory=pin),
                      torch.zeros(mx, dtype=torch.float64, device=device),
                      torch.zeros(world * mx, dtype=torch.float64, device=device),
                      torch.zeros(world * mx, dtype=torch.float64, pin_memory=pin))
    h_in, d_in, d_out, h_out = _bufs[key]
    m = np.asarray(mine, dtype=np.float64)
    if m.size:
        h_in.numpy()[:m.size] = m
    d_in.copy_(h_in, non_blocking=True)
    if hasattr(dist, "all_gather_into_tensor") and d_out.device.type == "cuda":
        dist.all_gather_into_tensor(d_out, d_in)
    else:   # gloo (CPU tests)
        parts = [d_out[r * mx:(r + 1) * mx] for r in range(world)]
        dist.all_gather(parts, d_in)
    h_out.copy_(d_out, non_blocking=True)
    if d_out.device.type == "cuda":
        torch.cuda.current_stream(d_out.device).synchronize()
    full = h_out.numpy()
    return np.concatenate([full[r * mx:r * mx + counts[r]] for r in range(world)])


def solve_items_sharded(hadi, num, pts, n, mode, local_solve, rank, world, device=None, dist=None):
    """Solve all items of a batch across `world` ranks.  local_solve(begin, end) -> values of that slice
    (on a GPU rank: a hadi Batch launch + fetch).  Returns every item value, identical on every rank."""
    costs = hadi.item_costs(num, pts, n, mode)
    sl = slices(hadi, costs, world)
    b, e = sl[rank]
    mine = local_solve(b, e) if e > b else np.zeros(0)
    if world == 1:
        return np.asarray(mine, dtype=np.float64)
    vpi = hadi.VALUES_PER_ITEM[mode]   # an item of the interpolated-V0 Jacobian publishes three values
    return allgather_values(mine, [(x[1] - x[0]) * vpi for x in sl], dist=dist, device=device)


def make_comm(hadi, rank, world, device=None, dist=None):
    """hadi_comm for hadi_calibrate: the C++ LM driver calls back into torch.distributed for its one
    all-gather per solver call.  Keep the returned object alive while it is in use."""
    def _cb(user, mine_p, my_count, all_p, counts_p, displs_p, w):
        try:
            counts = [counts_p[r] for r in range(w)]
            mine = np.ctypeslib.as_array(mine_p, shape=(max(my_count, 1),))[:my_count].copy()
            full = allgather_values(mine, counts, dist=dist, device=device)
            total = sum(counts)
            if total:
                np.ctypeslib.as_array(all_p, shape=(total,))[:] = full
            return 0
        except Exception:  # never let an exception cross the C ABI
            return 1

    fn = hadi.ALLGATHER_FN(_cb)
    comm = hadi.Comm(rank, world, fn, None)
    comm._keep = (fn, _cb)
    return comm


def attach_nccl(hadi, ctx, rank, world, dist=None, device=None):
    """Give `ctx` its own NCCL communicator (in-library exchange, include/hadi.h: hadi_comm_init): rank 0 creates the
    id, torch.distributed broadcasts the 128 bytes once, every rank joins.  After this hadi_calibrate(comm=None) and the
    *_sharded entry points run their all-gathers inside libhadi.so on the context's stream."""
    import torch

    if dist is None:
        import torch.distributed as dist
    t = torch.zeros(hadi.NCCL_ID_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(hadi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    ctx.comm_init(world, rank, bytes(t.cpu().numpy().tobytes()))
    return ctx
