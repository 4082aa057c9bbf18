"""Stress of the host-interaction paths of the sharded fit (stopping rules in batches, empty-cluster
relocation, NCCL fallback) against the single-GPU fit.  Run under torchrun."""
import os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-modal-regression_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
import torch.distributed as dist
from bdpose import kmeans
import test_gpu_multi as T
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
X = T._data()
n = X.shape[0]
lo, hi = rank * n // world, (rank + 1) * n // world
sh = lambda t: hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()[:10]
for case in T.CASES:
    init = X[:200].clone()
    kw = dict(max_iter=12)
    if case == "fixed":
        kw = dict(fixed_iters=5)
    if case == "empty":
        init[3] = torch.tensor([50.0, 50.0, 50.0], dtype=init.dtype)
    os.environ["BDPOSE_KMEANS_EXCHANGE"] = "p2p"
    os.environ["BDPOSE_KMEANS_NVLS"] = "0"
    ref = kmeans.kmeans_lloyd(X.cuda(), init.cuda(), group=kmeans.LOCAL, **kw)
    refs = [kmeans.kmeans_lloyd(X.cuda(), init.cuda(), group=kmeans.LOCAL, **kw) for _ in range(3)]
    self_ok = all(sh(r["centers"]) == sh(ref["centers"]) and r["n_iter"] == ref["n_iter"] for r in refs)
    bad, detail = 0, []
    for rep in range(reps):
        r = T._fit(kmeans, X, lo, hi, case)
        ok = sh(r["centers"]) == sh(ref["centers"]) and r["n_iter"] == ref["n_iter"] and \
            sh(r["labels"]) == sh(ref["labels"][lo:hi])
        t = torch.tensor([0 if ok else 1], device=dev)
        dist.all_reduce(t)
        if int(t) > 0:
            bad += 1
            detail.append((rep, r["n_iter"], ref["n_iter"], float((r["centers"] - ref["centers"]).abs().max())))
    if rank == 0:
        print("%s: single-GPU repeatable %s, n_iter %d; %d/%d sharded runs differ %s" % (
            case, self_ok, ref["n_iter"], bad, reps, detail), flush=True)
dist.barrier()
dist.destroy_process_group()
