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
tag = "dbg=%s" % os.environ.get("BDP_GEMM_DEBUG", "0")
B = 32
for precise in (False, True):
    out = []
    for N, K in ((16, 8192), (16, 16384), (256 * 148, 2048)):
        G = 1
        a = torch.randn(B, G * K, device=dev); w = torch.randn(G, N, K, device=dev); c = torch.empty(B, G * N, device=dev)
        us = t(lambda: head.gemm_tf32(a, 0, G * K, K, w, 0, K, N * K, c, 0, G * N, N, B, N, K, G=G, precise=precise))
        out.append("N=%d K=%d: %.1f" % (N, K, us))
    print(tag, "precise=%d" % precise, "  ".join(out))
