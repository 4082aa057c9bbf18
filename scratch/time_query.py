"""A/B of the query kernels: BDPOSE_QUERY_V1=1 selects the first-generation kernel.
usage: python scratch/time_query.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-modal-regression_b200"), ROOT]
import torch
from bench import synth_rotations, timed, kmeans_chunks
from bdpose import ops, kmeans, _lib as L
dev = torch.device("cuda", 0)
lib = L.lib()
x = synth_rotations(10_000_000, 1000, dev)
c = synth_rotations(1000, 7, dev).double().contiguous()
g = ops.KeyGrid(c)
for n in (10_000_000, 1_250_000):
    xs = x[:n].contiguous()
    ms = timed(lambda: ops.assign_nearest(xs, c, grid=g), 20, 5)
    print("assign f32 n=%d  %.1f us  %.0f GB/s" % (n, ms * 1e3, n * 32 / ms / 1e6))
    ms = timed(lambda: ops.assign_nearest(xs, c, grid=g, label_dtype=torch.int32, want_residual=False), 20, 5)
    print("assign f32 labels32 only n=%d  %.1f us" % (n, ms * 1e3))
xd = x.double().contiguous()
for n in (10_000_000, 5_000_000, 1_250_000):
    xs = xd[:n].contiguous()
    lab = torch.full((n,), -1, dtype=torch.int32, device=dev)
    A = 1000 * 7 + 2
    acc = torch.zeros(A, dtype=torch.int64, device=dev)
    def em():
        st = lib.bdp_kmeans_lloyd_step_grid(xs.data_ptr(), n, 3, c.data_ptr(), 1000, g.buf.data_ptr(), g.nbytes,
                                            lab.data_ptr(), acc.data_ptr(), 29, acc[A - 2:].data_ptr(), None, 1, L.stream_ptr())
        L.check(st, "lloyd")
    ms = timed(em, 20, 5)
    print("lloyd f64 n=%d  %.1f us  %.0f GB/s" % (n, ms * 1e3, n * 28 / ms / 1e6))
ms = timed(lambda: g.rebuild(), 20, 5)
print("grid build %.1f us" % (ms * 1e3))
