import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200"))
import torch
from bdpose import head
dev = torch.device("cuda", 0)
def t(fn, n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
B = 32
tag = "dbg=%s" % os.environ.get("BDP_GEMM_DEBUG", "0")
out = []
for precise in (False, True):
    for G, N, K in ((1, 96, 4096), (1, 176, 4096), (24, 500, 1000), (1, 24000, 2048)):
        a = torch.randn(B, G * K, device=dev); w = torch.randn(G, N, K, device=dev); c = torch.empty(B, G * N, device=dev)
        us = t(lambda: head.gemm_tf32(a, 0, G * K, K, w, 0, K, N * K, c, 0, G * N, N, B, N, K, G=G, precise=precise))
        out.append("%d/%dx%dx%d: %.1f" % (precise, G, N, K, us))
print(tag, "  ".join(out))
