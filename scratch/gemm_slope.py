import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200"))
import torch
from bdpose import head
dev = torch.device("cuda", 0)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
B = 32
for G, N in ((1, 96), (1, 176), (24, 576), (144, 96), (148, 176), (148, 256)):
    for K in (256, 512, 1024, 2048, 4096):
        a = torch.randn(B, G * K, device=dev); w = torch.randn(G, N, K, device=dev); c = torch.empty(B, G * N, device=dev)
        us = t(lambda: head.gemm_tf32(a, 0, G * K, K, w, 0, K, N * K, c, 0, G * N, N, B, N, K, G=G))
        print("G=%d N=%d K=%d: %.1f us  (%d kblocks, %.0f KB/tile, total %.1f MB, %.0f GB/s)" % (G, N, K, us, K // 32, (N + 32) * K * 4 / 1024, G * N * K * 4 / 1e6, G * N * K * 4 / us / 1e3))
