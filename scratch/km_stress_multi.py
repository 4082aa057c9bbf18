"""Stress of the sharded k-means paths: every configuration several times against the single-GPU
fit (run under torchrun).  usage: torchrun --nproc-per-node G scratch/km_stress_multi.py [n_total] [reps]"""
import os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-modal-regression_b200"), ROOT]
import torch
import torch.distributed as dist
from bench import kmeans_chunks, synth_rotations, N_ROT, N_CHUNKS
from bdpose import kmeans
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
X = kmeans_chunks(range(N_CHUNKS), dev)[:n].contiguous()
sh = lambda t: hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()[:10]
lo, hi = rank * n // world, (rank + 1) * n // world
for K in (200, 1000):
    init = synth_rotations(N_ROT // N_CHUNKS, 100, dev, torch.float64)[:K].clone()
    ref = kmeans.kmeans_lloyd(X, init, fixed_iters=12, group=kmeans.LOCAL)
    href, hlab = sh(ref["centers"]), sh(ref["labels"][lo:hi])
    for shard in ("1", "0"):
        for inc in ("1", "0"):
            for nvls in ("0", "1"):
                os.environ["BDPOSE_KMEANS_SHARD_GRID"] = shard
                os.environ["BDPOSE_KMEANS_INCREMENTAL"] = inc
                os.environ["BDPOSE_KMEANS_NVLS"] = nvls
                bad = 0
                for rep in range(reps):
                    kmeans.Exchange._cache.clear()
                    r = kmeans.kmeans_lloyd(X[lo:hi], init, fixed_iters=12)
                    ok = sh(r["centers"]) == href and sh(r["labels"]) == hlab
                    t = torch.tensor([0 if ok else 1], device=dev)
                    dist.all_reduce(t)
                    bad += int(t) > 0
                if rank == 0:
                    print("K=%d shard=%s incremental=%s nvls=%s: %d/%d runs differ (%s)" % (K, shard, inc, nvls, bad, reps, r["exchange"]), flush=True)
dist.barrier()
dist.destroy_process_group()
