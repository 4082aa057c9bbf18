import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200"))
import torch
import bench_head
from bdpose import head, ops, _lib as L
dev = torch.device("cuda", 0)
m = bench_head._pascal_model().train()
keys = torch.randn(200, 3, device=dev)
B = 32
x = torch.randn(B, 2048, device=dev, requires_grad=True)
lab = torch.randint(0, 12, (B, 1), device=dev)
bins = torch.randint(0, 200, (B,), device=dev)
tgt = torch.randn(B, 3, device=dev)
head.set_precision("tf32")
def step():
    for p in m.parameters():
        p.grad = None
    y1, y2 = m(x, lab)
    lc, lr, _ = ops.bd_loss(y1, bins, y2, tgt, keys, L.POSE_GEODESIC_AA, True)
    (lc + lr).backward()
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(50): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
