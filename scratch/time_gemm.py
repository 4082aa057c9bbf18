import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200"))
import torch
from bdpose import head
dev = torch.device("cuda", 0)
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
H, N1, N2, N0 = 24, 1000, 500, 2048
F1, F2 = H * N1, H * N2
for B in (32, 96):
    x = torch.randn(B, N0, device=dev); w1 = torch.randn(F1, N0, device=dev); h1 = torch.empty(B, F1, device=dev)
    w2 = torch.randn(H, N2, N1, device=dev); h2 = torch.empty(B, F2, device=dev)
    dw1 = torch.empty(F1, N0, device=dev); dw2 = torch.empty(H, N2, N1, device=dev)
    for precise in (False, True):
        us = t(lambda: head.gemm_tf32(x, 0, N0, 0, w1, 0, N0, 0, h1, 0, F1, 0, B, F1, N0, precise=precise))
        print("B=%d fc1 precise=%d: %.1f us  %.0f GB/s" % (B, precise, us, F1 * N0 * 4 / us / 1e3))
        us = t(lambda: head.gemm_tf32(h1, 0, F1, N1, w2, 0, N1, N2 * N1, h2, 0, F2, N2, B, N2, N1, G=H, precise=precise))
        print("B=%d fc2 precise=%d: %.1f us %.0f GB/s" % (B, precise, us, H * N2 * N1 * 4 / us / 1e3))
        us = t(lambda: head.gemm_tf32(h2, 0, F2, N2, w2, 1, N1, N2 * N1, h1, 0, F1, N1, B, N1, N2, G=H, precise=precise))
        print("B=%d fc2 dgrad precise=%d: %.1f us %.0f GB/s" % (B, precise, us, H * N2 * N1 * 4 / us / 1e3))
        us = t(lambda: head.gemm_tf32(h2, 1, F2, N2, h1, 1, F1, N1, dw2, 0, N1, N2 * N1, N2, N1, B, G=H, precise=precise))
        print("B=%d fc2 wgrad precise=%d: %.1f us %.0f GB/s" % (B, precise, us, H * N2 * N1 * 4 / us / 1e3))
        us = t(lambda: head.gemm_tf32(h1, 1, F1, 0, x, 1, N0, 0, dw1, 0, N0, 0, F1, N0, B, precise=precise))
        print("B=%d fc1 wgrad precise=%d: %.1f us %.0f GB/s" % (B, precise, us, F1 * N0 * 4 / us / 1e3))
        S = head.gemm_splits(F1, 18)
        parts = torch.empty(S, B, N0, device=dev)
        us = t(lambda: head.gemm_tf32(h1, 0, F1, 0, w1, 1, N0, 0, parts, 0, N0, 0, B, N0, F1, splits=18, c_ss=B * N0, precise=precise))
        print("B=%d fc1 dgrad precise=%d splits=%d: %.1f us %.0f GB/s" % (B, precise, S, us, F1 * N0 * 4 / us / 1e3))
