"""world_size-2 gloo tests (CPU) of the multi-rank host logic: the k-means host loop of
bdpose.kmeans.kmeans_lloyd (row sharding, one all-reduce of the int64 accumulators per iteration,
convergence tests, empty-cluster relocation with its all-gather) and the head-gradient all-reduce.
The CUDA entry points are replaced by the numpy stand-ins of oracle/lloyd_host.py (same C-ABI
contracts); everything else is the product's own code."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _paths():
    for p in (os.path.join(ROOT, "multi-modal-regression_b200"), os.path.join(ROOT, "oracle"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)


def _rot(rng, n):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 0] < 0] *= -1
    ang = 2 * np.arccos(np.clip(q[:, 0], -1, 1))
    ax = q[:, 1:] / np.maximum(np.linalg.norm(q[:, 1:], axis=1, keepdims=True), 1e-300)
    return ax * ang[:, None]


def _problem(case):
    rng = np.random.default_rng(5)
    X = _rot(rng, 6000)
    K = 24
    init = X[:K].copy()
    if case == "empty":
        init[3] = init[4] + 1e-9        # two seeds on top of each other -> one cluster starves
        init[7] = [40.0, 40.0, 40.0]    # a seed far away from everything: empty from iteration 1
    return X, init


def _kmeans_worker(rank, world, store, case, out):
    _paths()
    dist.init_process_group("gloo", init_method="file://" + store, rank=rank, world_size=world)
    from bdpose import kmeans
    import lloyd_host
    X, init = _problem(case)
    cut = [0, 2500, 6000] if world == 2 else [0, 6000]       # deliberately uneven shards
    xs = torch.from_numpy(X[cut[rank]:cut[rank + 1]].copy())
    r = kmeans.kmeans_lloyd(xs, torch.from_numpy(init), max_iter=30, _backend=lloyd_host.BACKEND)
    torch.save({"centers": r["centers"], "labels": r["labels"], "n_iter": r["n_iter"],
                "inertia": r["inertia"]}, os.path.join(out, "km_%s_%d_%d.pt" % (case, world, rank)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["plain", "empty"])
def test_kmeans_host_loop_two_ranks_equals_one(case):
    _paths()
    import bdpose_oracle as O
    with tempfile.TemporaryDirectory() as out:
        for world in (1, 2):
            store = os.path.join(out, "store_%d" % world)
            mp.spawn(_kmeans_worker, args=(world, store, case, out), nprocs=world, join=True)
        one = torch.load(os.path.join(out, "km_%s_1_0.pt" % case))
        two = [torch.load(os.path.join(out, "km_%s_2_%d.pt" % (case, r))) for r in range(2)]
    # every rank ends with the same centres, identical to the single-rank run bit for bit
    assert torch.equal(two[0]["centers"], two[1]["centers"])
    assert torch.equal(two[0]["centers"], one["centers"])
    assert two[0]["n_iter"] == two[1]["n_iter"] == one["n_iter"]
    assert torch.equal(torch.cat([two[0]["labels"], two[1]["labels"]]), one["labels"])
    assert abs(two[0]["inertia"] - one["inertia"]) <= 1e-9 * one["inertia"]
    if case == "plain":
        X, init = _problem(case)
        ref = O.kmeans_lloyd(X, init, max_iter=30)
        assert np.array_equal(one["labels"].numpy(), ref["labels"])
        assert np.allclose(one["centers"].numpy(), ref["centers"], rtol=0, atol=1e-12)
        assert one["n_iter"] == ref["n_iter"]


def _grad_worker(rank, world, store, out):
    _paths()
    dist.init_process_group("gloo", init_method="file://" + store, rank=rank, world_size=world)
    from bdpose import head
    torch.manual_seed(0)
    mods = [torch.nn.Sequential() for _ in range(2)]
    stack = head.HeadStack.__new__(head.HeadStack)
    # stacked gradient buffers as _HeadFn.backward deposits them (rank-dependent values)
    g = torch.Generator().manual_seed(1)
    base = {"w1": torch.randn(4, 6, 8, generator=g), "g1": torch.randn(4, 6, generator=g),
            "w3_0": torch.randn(2, 5, 3, generator=g), "b3_0": torch.randn(2, 5, generator=g)}
    stack.grad = {k: v * (rank + 1) for k, v in base.items()}
    head.allreduce_stack_grads(stack, group=None)
    mean = sum(range(1, world + 1)) / world
    for k, v in base.items():
        assert torch.allclose(stack.grad[k], v * mean, rtol=1e-6, atol=1e-7), k
    # Parameters whose .grad are NOT views of the stacked buffers (foreign gradients were present
    # when backward ran): the gradients the Parameters hold are the ones reduced
    stack2 = head.HeadStack.__new__(head.HeadStack)
    ps = [torch.nn.Parameter(torch.zeros(3, 2)) for _ in range(3)]
    for i, p_ in enumerate(ps):
        p_.grad = torch.full((3, 2), float((rank + 1) * (i + 1)))
    stack2.grad = {"w1": torch.zeros(3, 3, 2)}              # stale stacked buffer: must stay untouched
    stack2.gflat, stack2.stacked = None, None
    stack2.plists = {"w1": ps}
    head.allreduce_stack_grads(stack2, group=None)
    for i, p_ in enumerate(ps):
        assert torch.allclose(p_.grad, torch.full((3, 2), mean * (i + 1))), i
    assert float(stack2.grad["w1"].abs().max()) == 0.0
    torch.save({"ok": True}, os.path.join(out, "g_%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_head_gradient_allreduce_two_ranks():
    with tempfile.TemporaryDirectory() as out:
        mp.spawn(_grad_worker, args=(2, os.path.join(out, "store"), out), nprocs=2, join=True)
        assert all(os.path.exists(os.path.join(out, "g_%d.pt" % r)) for r in range(2))
