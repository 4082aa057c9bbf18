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


CASES = ("converge", "fixed", "nccl", "empty", "nvls")


def _fit(kmeans, X, lo, hi, case):
    init = X[:200].clone()
    kw = dict(max_iter=12)
    if case == "fixed":
        kw = dict(fixed_iters=5)
    if case == "empty":
        init[3] = torch.tensor([50.0, 50.0, 50.0], dtype=init.dtype)   # a centre no sample is near
    os.environ["BDPOSE_KMEANS_EXCHANGE"] = "nccl" if case == "nccl" else "p2p"
    os.environ["BDPOSE_KMEANS_NVLS"] = "1" if case == "nvls" else "0"
    kmeans.Exchange._cache.clear()
    return kmeans.kmeans_lloyd(X[lo:hi].cuda(), init.cuda(), **kw)


def _worker(rank, world, store, out):
    _paths()
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="file://" + store, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    from bdpose import kmeans
    X = _data()
    n = X.shape[0]
    lo, hi = rank * n // world, (rank + 1) * n // world
    res = {}
    for case in CASES:
        r = _fit(kmeans, X, lo, hi, case)
        res[case] = {"centers": r["centers"].cpu(), "labels": r["labels"].cpu(), "n_iter": r["n_iter"],
                     "inertia": r["inertia"], "exchange": r["exchange"]}
    torch.save(res, os.path.join(out, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_kmeans_equals_single_gpu():
    """Sharded fits (fused peer-memory exchange, its NVLS form, the NCCL fallback; stopping rules,
    fixed iterations, an empty-cluster relocation) against the single-GPU fit of the same data:
    centres, labels and iteration counts bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _paths()
    from bdpose import kmeans
    world = min(torch.cuda.device_count(), 8)
    with tempfile.TemporaryDirectory() as out:
        mp.spawn(_worker, args=(world, os.path.join(out, "store"), out), nprocs=world, join=True)
        parts = [torch.load(os.path.join(out, "r%d.pt" % r)) for r in range(world)]
    X = _data()
    modes = {c: parts[0][c]["exchange"] for c in CASES}
    print("exchange modes:", modes)
    assert modes["nccl"] == "nccl"
    assert modes["converge"] in ("p2p", "nccl"), modes      # nccl only where symmetric memory is absent
    bad = []
    for case in CASES:
        one = _fit(kmeans, X, 0, X.shape[0], case)
        for r, p in enumerate(parts):
            if p[case]["n_iter"] != one["n_iter"]:
                bad.append("%s: rank %d ran %d iterations, single GPU %d" % (case, r, p[case]["n_iter"], one["n_iter"]))
            elif not torch.equal(p[case]["centers"], one["centers"].cpu()):
                bad.append("%s: centres of rank %d differ (max %.3e)" % (
                    case, r, float((p[case]["centers"] - one["centers"].cpu()).abs().max())))
            elif p[case]["inertia"] != pytest.approx(one["inertia"], rel=1e-12):
                bad.append("%s: inertia of rank %d" % (case, r))
        if not torch.equal(torch.cat([p[case]["labels"] for p in parts]), one["labels"].cpu()):
            bad.append("%s: labels differ" % case)
    assert not bad, bad
