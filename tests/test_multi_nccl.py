"""In-library multi-GPU exchange (hadi_comm_init / *_sharded / hadi_calibrate over the context's NCCL communicator)
on real GPUs: every rank must reproduce the single-GPU numbers bit for bit.  Needs at least two GPUs
(gpurun --gpus 2); skipped elsewhere.  The host-side sharding logic is covered on CPU by test_multi_gloo.py."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_nccl_sharded_entry_points_equal_single_gpu():
    import torch

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29713", os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [json.loads(ln.split("NCCLWORKER ", 1)[1]) for ln in r.stdout.splitlines() if "NCCLWORKER " in ln]
    assert r.returncode == 0 and len(lines) == world, r.stdout[-2000:] + r.stderr[-2000:]
    for d in lines:
        bad = [k for k, v in d.items() if k != "rank" and v is not True]
        assert not bad, (d["rank"], bad)
