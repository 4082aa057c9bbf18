import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200"))
sys.path.insert(0, ROOT)
import torch
from bdpose import ops, kmeans
from bench import synth_rotations
dev = torch.device("cuda", 0)
N = 10_000_000
x = synth_rotations(N, 1000, dev)
xd = x.double()
K = 1000
c = synth_rotations(K, 7, dev).double().contiguous()
g = ops.KeyGrid(c)
for _ in range(2):
    ops.assign_nearest(x, c, grid=g)
hb = kmeans._fix_hi_bits(float(xd.abs().max()))
st = kmeans.LloydState(N, K, 3, dev)
for upd in (True, False, True):
    st.acc_stats.zero_()
    kmeans.lloyd_step(xd, c, st, hb, update=upd, grid=g)
torch.cuda.synchronize()
print("done")
