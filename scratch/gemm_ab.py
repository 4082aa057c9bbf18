"""Same-process A/B of two builds of bdp_gemm_tf32 on the same buffers."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200"))
import torch
from bdpose import _lib as L
dev = torch.device("cuda", 0)
new = L.lib()
import glob
sig = L.SIGNATURES["bdp_gemm_tf32"]
libs = []
for f in sorted(glob.glob(os.path.join(ROOT, "scratch", "libgemm_*.so")), key=os.path.getmtime):
    l = C.CDLL(f)
    l.bdp_gemm_tf32.restype, l.bdp_gemm_tf32.argtypes = sig[0], sig[1]
    libs.append((os.path.basename(f)[8:-3], l))
libs.append(("new", new))
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
H, N1, N2, N0, B = 24, 1000, 500, 2048, 32
F1, F2 = H * N1, H * N2
x = torch.randn(B, N0, device=dev); w1 = torch.randn(F1, N0, device=dev); h1 = torch.empty(B, F1, device=dev)
w2 = torch.randn(H, N2, N1, device=dev); h2 = torch.empty(B, F2, device=dev)
st = L.stream_ptr()
for rep in range(2):
    for name, lib in libs:
        for precise in (0, 1):
            a = t(lambda: lib.bdp_gemm_tf32(x.data_ptr(), 0, N0, 0, w1.data_ptr(), 0, N0, 0, h1.data_ptr(), 0, F1, 0, B, F1, N0, 1, 1, 0, precise, st))
            b = t(lambda: lib.bdp_gemm_tf32(h1.data_ptr(), 0, F1, N1, w2.data_ptr(), 0, N1, N2 * N1, h2.data_ptr(), 0, F2, N2, B, N2, N1, H, 1, 0, precise, st))
            print("%s precise=%d fc1 %.1f us  fc2 %.1f us" % (name, precise, a, b))
