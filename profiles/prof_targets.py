"""Short driver for ncu: runs each hot kernel a few times on BASELINE-sized inputs.
usage: python profiles/prof_targets.py [assign|loss|lloyd|eval|head ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "multi-modal-regression_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402
from bench import synth_rotations, N_ROT, K_DICT  # noqa: E402
from bdpose import ops, kmeans, _lib as L  # noqa: E402

which = sys.argv[1:] or ["assign", "loss", "lloyd", "eval"]
dev = torch.device("cuda", 0)
x = synth_rotations(N_ROT, 1000, dev)
centers = synth_rotations(K_DICT, 7, dev).double().contiguous()
reps = 3
if "assign" in which:
    for _ in range(reps):
        ops.assign_nearest(x, centers)
if "loss" in which:
    B, K = 1_000_000, 200
    score = torch.randn(B, K, device=dev)
    bins = torch.randint(0, K, (B,), device=dev)
    delta = torch.randn(B, 3, device=dev) * 0.2
    keys = centers[:K].float().contiguous()
    for _ in range(reps):
        ops.bd_loss_raw(score, bins, delta, x[:B].contiguous(), keys, L.POSE_GEODESIC_AA, True)
if "lloyd" in which:
    xd = x.double().contiguous()
    kmeans.kmeans_lloyd(xd, centers.clone(), fixed_iters=8, group=kmeans.LOCAL)
if "eval" in which:
    a = x[:1_000_000].contiguous(); b = x[1_000_000:2_000_000].contiguous()
    labels = torch.randint(0, 12, (1_000_000,), device=dev)
    for _ in range(reps):
        e = ops.geodesic_error_deg(a, b)
        ops.error_stats(e, labels, 12)
if "head" in which:
    import bench_head
    bench_head.profile(dev)
torch.cuda.synchronize()
print("done")
