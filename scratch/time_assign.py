import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200"))
sys.path.insert(0, ROOT)
import torch
from bdpose import ops, kmeans
from bench import synth_rotations
dev = torch.device("cuda", 0)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
N = 10_000_000
x = synth_rotations(N, 1000, dev)
xd = x.double()
for K in (1000, 200, 16):
    c = synth_rotations(K, 7, dev).double().contiguous()
    g = ops.KeyGrid(c)
    print("K=%d grid bytes %.1f MB" % (K, g.nbytes / 1e6))
    print("  build            %.1f us" % t(lambda: g.rebuild()))
    print("  assign grid      %.1f us" % t(lambda: ops.assign_nearest(x, c, grid=g)))
    print("  assign auto      %.1f us" % t(lambda: ops.assign_nearest(x, c)))
    print("  assign brute     %.1f us" % t(lambda: ops.assign_nearest(x, c, grid=None), 3))
    hb = kmeans._fix_hi_bits(float(xd.abs().max()))
    st = kmeans.LloydState(N, K, 3, dev)
    def step(grid):
        st.acc_stats.zero_()
        kmeans.lloyd_step(xd, c, st, hb, update=True, grid=grid)
    print("  lloyd grid       %.1f us" % t(lambda: step(g)))
    print("  lloyd brute      %.1f us" % t(lambda: step(None), 3))
    # statistics of the candidate lists
    import numpy as np
    buf = g.buf.cpu().numpy()
    G = int(np.frombuffer(buf[96:100].tobytes(), dtype=np.int32)[0])
    ncoarse = (G // 4) ** 3
    fine = np.frombuffer(buf[160:160 + G ** 3 * 64].tobytes(), dtype=np.uint16).reshape(-1, 32)
    cnt = fine[:, 0].astype(np.int64)
    print("  G=%d cells=%d mean cand %.2f max %d overflow cells %d  hist %s" % (G, G ** 3, cnt[cnt < 65535].mean(), cnt[cnt < 65535].max(), int((cnt == 65535).sum()), np.bincount(cnt[cnt < 65535])[:12].tolist()))
