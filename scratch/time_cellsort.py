"""Cost of bdp_cellsort's pieces (10 M rotations): key kernel + CUB radix sort + row gather, and the
label scatter.  usage: python scratch/time_cellsort.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-modal-regression_b200"), ROOT]
import torch
from bench import kmeans_chunks, synth_rotations, timed, N_ROT, N_CHUNKS, K_DICT
from bdpose import ops, kmeans, _lib as L
dev = torch.device("cuda", 0)
lib = L.lib()
xs = kmeans_chunks(range(N_CHUNKS), dev)
init = synth_rotations(N_ROT // N_CHUNKS, 100, dev, torch.float64)[:K_DICT].clone()
fs = kmeans.FitSetup(xs, init, group=kmeans.LOCAL)
x, N, d, K = fs.x, fs.x.shape[0], 3, K_DICT
g = ops.KeyGrid(fs.centers, build=False)
lo, hi = x.new_full((3,), -fs.max_abs), x.new_full((3,), fs.max_abs)
L.check(lib.bdp_keygrid_prepare(lo.data_ptr(), hi.data_ptr(), K, d, g.buf.data_ptr(), g.nbytes, L.stream_ptr()), "prep")
nws = lib.bdp_cellsort_workspace_bytes(N, K, d)
ws = torch.empty(nws, dtype=torch.uint8, device=dev)
perm = torch.empty(N, dtype=torch.int32, device=dev)
out = torch.empty_like(x)
occ = torch.zeros(lib.bdp_keygrid_coarse_cells(K, d), dtype=torch.int32, device=dev)
ms = timed(lambda: L.check(lib.bdp_cellsort(x.data_ptr(), N, d, K, g.buf.data_ptr(), g.nbytes, occ.data_ptr(),
                                            ws.data_ptr(), nws, perm.data_ptr(), out.data_ptr(), L.stream_ptr()), "s"), 10, 3)
print("bdp_cellsort (keys + sort + gather): %.1f us, workspace %.0f MB" % (ms * 1e3, nws / 1e6))
ms = timed(lambda: L.check(lib.bdp_keygrid_occupancy(x.data_ptr(), N, d, K, g.buf.data_ptr(), g.nbytes,
                                                     occ.data_ptr(), L.stream_ptr()), "o"), 10, 3)
print("occupancy pass alone (~ key kernel): %.1f us" % (ms * 1e3))
lab = torch.zeros(N, dtype=torch.int32, device=dev)
dst = torch.zeros(N, dtype=torch.int32, device=dev)
ms = timed(lambda: L.check(lib.bdp_scatter_i32(lab.data_ptr(), perm.data_ptr(), N, dst.data_ptr(), L.stream_ptr()), "sc"), 10, 3)
print("label scatter: %.1f us" % (ms * 1e3))
p = perm.long()
ms = timed(lambda: torch.index_select(x, 0, p), 10, 3)
print("torch row gather (index_select): %.1f us" % (ms * 1e3))
