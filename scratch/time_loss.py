import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200")); sys.path.insert(0, ROOT)
import torch
from bdpose import ops, _lib as L
from bench import synth_rotations
dev = torch.device("cuda", 0)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
B, K = 1_000_000, 200
score = torch.randn(B, K, device=dev)
bins = torch.randint(0, K, (B,), device=dev)
delta = torch.randn(B, 3, device=dev) * 0.2
target = synth_rotations(B, 3, dev)
keys = synth_rotations(K, 7, dev)
for mode, name in ((L.POSE_GEODESIC_AA, "geodesic aa"), (L.POSE_MSE, "mse")):
    us = t(lambda: ops.bd_loss_raw(score, bins, delta, target, keys, mode, True))
    by = B * (2 * K * 4 + 8 + 12 + 12 + 12 + 8)
    print("%s: %.1f us  %.0f GB/s (%.1f%% of 6551)" % (name, us, by / us / 1e3, by / us / 1e3 / 65.51))
us = t(lambda: ops.bd_loss_raw(score, bins, delta, target, keys, L.POSE_GEODESIC_AA, True, want_grad=False))
print("fwd only: %.1f us" % us)
