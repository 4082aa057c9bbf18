import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200"))
import torch
from bdpose import head
dev = torch.device("cuda", 0)
H, N1, N0, B = 24, 1000, 2048, 32
w1 = torch.randn(H * N1, N0, device=dev)
x = torch.randn(B, N0, device=dev)
h1 = torch.empty(H * N1, 32, device=dev)
for precise in (False, True):
    for _ in range(3):
        head.gemm_tf32(w1, 0, N0, 0, x, 0, N0, 0, h1, 0, 32, 0, H * N1, B, N0, precise=precise)
# wgrad shape: dW1 = dH1^T x X
dh = torch.randn(H * N1, 32, device=dev)
dw = torch.empty(H * N1, N0, device=dev)
for _ in range(3):
    head.gemm_tf32(dh, 0, 32, 0, x, 1, N0, 0, dw, 0, N0, 0, H * N1, N0, B, precise=True)
torch.cuda.synchronize()
print("done")
