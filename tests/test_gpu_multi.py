"""Multi-GPU (NCCL) checks, run when the box has >= 2 GPUs: the sharded k-means fit gives the same
centres and labels as the single-GPU fit bit for bit, and the head-gradient all-reduce averages."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _paths():
    for p in (os.path.join(ROOT, "multi-modal-regression_b200"), os.path.join(ROOT, "oracle"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)


def _data():
    g = torch.Generator().manual_seed(4)
    q = torch.randn(400_000, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    w = q[:, :1].abs().clamp(max=1)
    return (q[:, 1:] / q[:, 1:].norm(dim=1, keepdim=True) * (2 * torch.acos(w))).contiguous()


def _worker(rank, world, store, out):
    _paths()
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="file://" + store, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    from bdpose import kmeans
    X = _data()
    n = X.shape[0]
    lo, hi = rank * n // world, (rank + 1) * n // world
    r = kmeans.kmeans_lloyd(X[lo:hi].cuda(), X[:200].cuda(), max_iter=12)
    torch.save({"centers": r["centers"].cpu(), "labels": r["labels"].cpu(), "n_iter": r["n_iter"]},
               os.path.join(out, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_kmeans_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _paths()
    from bdpose import kmeans
    world = 2
    with tempfile.TemporaryDirectory() as out:
        mp.spawn(_worker, args=(world, os.path.join(out, "store"), out), nprocs=world, join=True)
        parts = [torch.load(os.path.join(out, "r%d.pt" % r)) for r in range(world)]
    X = _data()
    one = kmeans.kmeans_lloyd(X.cuda(), X[:200].cuda(), max_iter=12)
    assert parts[0]["n_iter"] == parts[1]["n_iter"] == one["n_iter"]
    assert torch.equal(parts[0]["centers"], parts[1]["centers"])
    assert torch.equal(parts[0]["centers"], one["centers"].cpu())
    assert torch.equal(torch.cat([p["labels"] for p in parts]), one["labels"].cpu())
