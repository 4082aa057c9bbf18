"""Per-iteration anatomy of the k-means loop on one GPU: E+M kernel times of every iteration (events
inside bdp_kmeans_run), whole-loop time per iteration, grid build alone.
usage: python scratch/time_kmeans.py [n_rotations]      (env BDPOSE_QUERY_THREADS=512 to A/B)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-modal-regression_b200"), ROOT]
import torch
from bench import kmeans_chunks, synth_rotations, timed, N_ROT, N_CHUNKS, K_DICT
from bdpose import ops, kmeans, _lib as L
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else N_ROT
xs = kmeans_chunks(range(N_CHUNKS), dev)[:n].contiguous()
init = synth_rotations(N_ROT // N_CHUNKS, 100, dev, torch.float64)[:K_DICT].clone()
fs = kmeans.FitSetup(xs, init, group=kmeans.LOCAL)
labels = torch.full((n,), -1, dtype=torch.int32, device=dev)
grid = ops.KeyGrid(fs.centers)
loop = kmeans.LloydLoop(fs.x, fs.centers, labels, fs.hb, grid, kmeans.LOCAL, fs.tol_abs)
steps = 20
for rep in range(2):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4 * steps)]
    for e in evs:
        e.record()
    loop.reset(fs.centers)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop.launch(0, steps, False, em_events=evs)
    e1.record()
    torch.cuda.synchronize()
em = [evs[4 * i + 1].elapsed_time(evs[4 * i + 2]) * 1e3 for i in range(steps)]
print('  build us', ' '.join('%.0f' % (evs[4*i].elapsed_time(evs[4*i+1])*1e3) for i in range(steps)))
print('  xfin  us', ' '.join('%.0f' % (evs[4*i+2].elapsed_time(evs[4*i+3])*1e3) for i in range(steps)))
print("n=%d incremental=%s: loop %.1f us/iter; E+M kernel us per iteration: %s" % (
    n, loop.incremental, e0.elapsed_time(e1) * 1e3 / steps, " ".join("%.0f" % v for v in em)))
print("  E+M mean %.1f us -> %.0f GB/s (28 B/rot)" % (sum(em) / steps, n * 28 / (sum(em) / steps) / 1e3))
print("  grid build alone %.1f us" % (timed(lambda: grid.rebuild(), 20, 5) * 1e3))
st = loop.status()
print("  status", st.state, st.iter_done, st.changed, st.shift2)
