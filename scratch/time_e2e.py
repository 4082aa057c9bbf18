import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200")); sys.path.insert(0, ROOT)
import torch
import binDeltaGenerators as G
from bench import synth_rotations
dev = torch.device("cuda", 0)
N = 10_000_000
x = synth_rotations(N, 1000, dev)
c = synth_rotations(1000, 7, dev).double().contiguous()
y = x.cpu().pin_memory()
ob = torch.empty(N, dtype=torch.int64).pin_memory()
orr = torch.empty(N, 3, dtype=torch.float32).pin_memory()
print("pinned:", y.is_pinned(), y[5:100].is_pinned(), ob[5:100].is_pinned())
for chunk in (10_000_000, 2_500_000, 1 << 20, 1 << 19):
    G.assign_labels_host(y, c, ob, orr, chunk_rows=chunk)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        G.assign_labels_host(y, c, ob, orr, chunk_rows=chunk)
    torch.cuda.synchronize()
    print("chunk %8d: %.2f ms" % (chunk, (time.perf_counter() - t0) / 5 * 1e3))
