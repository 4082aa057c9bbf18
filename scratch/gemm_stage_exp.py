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
tag = "stages=%s rot=%s" % (os.environ.get("BDP_GEMM_STAGES", "max"), os.environ.get("BDP_GEMM_NO_ROTATE", "0"))
for G, N, K, shared_a in ((1, 96, 4096, 0), (1, 24000, 2048, 0), (148, 176, 2048, 0), (148, 176, 2048, 1), (24, 500, 1000, 0), (24, 500, 1000, 1)):
    a = torch.randn(B, K if shared_a else G * K, device=dev); w = torch.randn(G, N, K, device=dev); c = torch.empty(B, G * N, device=dev)
    if shared_a:
        us = t(lambda: head.gemm_tf32(a, 0, K, 0, w, 0, K, N * K, c, 0, G * N, N, B, N, K, G=G))
    else:
        us = t(lambda: head.gemm_tf32(a, 0, G * K, K, w, 0, K, N * K, c, 0, G * N, N, B, N, K, G=G))
    print("%s G=%d N=%d K=%d sharedA=%d: %.1f us (%.0f GB/s)" % (tag, G, N, K, shared_a, us, G * N * K * 4 / us / 1e3))
