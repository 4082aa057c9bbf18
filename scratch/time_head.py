import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200"))
import torch
import bench_head
from bdpose import head, ops, _lib as L
dev = torch.device("cuda", 0)
m = bench_head._pascal_model().train()
keys = torch.randn(200, 3, device=dev)
params = list(m.parameters())
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3, (t1 - t0) / n * 1e6
for B in (32, 96):
    x = torch.randn(B, 2048, device=dev, requires_grad=True)
    lab = torch.randint(0, 12, (B, 1), device=dev)
    bins = torch.randint(0, 200, (B,), device=dev)
    tgt = torch.randn(B, 3, device=dev)
    def fwd():
        with torch.no_grad():
            m(x, lab)
    def step():
        for p in params:
            p.grad = None
        y1, y2 = m(x, lab)
        lc, lr, _ = ops.bd_loss(y1, bins, y2, tgt, keys, L.POSE_GEODESIC_AA, True)
        (lc + lr).backward()
    for mode in ("tf32", "fp32"):
        head.set_precision(mode)
        g, h = t(fwd)
        print("B=%d %s fwd      gpu %.0f us  host-issue %.0f us" % (B, mode, g, h))
        g, h = t(step)
        print("B=%d %s fwd+bwd  gpu %.0f us  host-issue %.0f us" % (B, mode, g, h))
    head.set_precision("fp32")

# opt-in fast path: stacked parameters (10 tensors instead of 336)
sp = m.stacked_head_parameters()
for B in (32, 96):
    x = torch.randn(B, 2048, device=dev, requires_grad=True)
    lab = torch.randint(0, 12, (B, 1), device=dev)
    bins = torch.randint(0, 200, (B,), device=dev)
    tgt = torch.randn(B, 3, device=dev)
    def step2():
        for p in sp:
            p.grad = None
        y1, y2 = m(x, lab)
        lc, lr, _ = ops.bd_loss(y1, bins, y2, tgt, keys, L.POSE_GEODESIC_AA, True)
        (lc + lr).backward()
    for mode in ("tf32", "fp32"):
        head.set_precision(mode)
        g, h = t(step2)
        print("B=%d %s fwd+bwd STACKED gpu %.0f us  host-issue %.0f us" % (B, mode, g, h))

# CUDA-graph step
from bdpose.graph_step import GraphedBinDeltaStep
for B in (32, 96):
    x = torch.randn(B, 2048, device=dev)
    lab = torch.randint(0, 12, (B, 1), device=dev)
    bins = torch.randint(0, 200, (B,), device=dev)
    tgt = torch.randn(B, 3, device=dev)
    for mode in ("tf32", "fp32"):
        head.set_precision(mode)
        gs = GraphedBinDeltaStep(m, B, keys, L.POSE_GEODESIC_AA, True)
        g, h = t(lambda: gs(x, lab, bins, tgt))
        print("B=%d %s fwd+bwd GRAPH   gpu %.0f us  host-issue %.0f us" % (B, mode, g, h))
head.set_precision("fp32")
