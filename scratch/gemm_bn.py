import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200"))
import torch
from bdpose import head
dev = torch.device("cuda", 0)
def t(fn, n=30):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
tag = "dbg=%s BN=%s" % (os.environ.get("BDP_GEMM_DEBUG", "0"), os.environ.get("BDP_GEMM_BN"))
B = 32
N = int(os.environ["BDP_GEMM_BN"])
out = []
for precise in (False, True):
    ts = []
    for K in (8192, 16384):
        a = torch.randn(B, K, device=dev); w = torch.randn(1, N, K, device=dev); c = torch.empty(B, N, device=dev)
        ts.append(t(lambda: head.gemm_tf32(a, 0, K, K, w, 0, K, N * K, c, 0, N, N, B, N, K, G=1, precise=precise)))
    out.append("precise=%d: %.3f us/kblock" % (precise, (ts[1] - ts[0]) / 256))
print(tag, "  ".join(out))
