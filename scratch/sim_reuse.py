"""CPU model of key-grid list reuse across Lloyd iterations (no GPU): per-iteration max centre shift of the
bench workload (uniform rotations, K = 1000, init = first K rows), how often a build with slack
delta = alpha * max shift stays valid, and what the slack does to the candidate lists.
usage: python scratch/sim_reuse.py [N] [iters]"""
import sys
import numpy as np
import torch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 23
K, G = 1000, 64
torch.manual_seed(0)
q = torch.randn(N, 4, dtype=torch.float64)
q /= q.norm(dim=1, keepdim=True)
w = q[:, :1].abs().clamp(max=1)
v = q[:, 1:] * torch.sign(q[:, :1])
x = (v / v.norm(dim=1, keepdim=True).clamp_min(1e-30) * (2 * torch.acos(w))).float()
x -= x.mean(0)
c = x[:K].clone().double()
box = float(x.abs().max())
cell = 2 * box / G
print("box %.4f cell %.4f" % (box, cell))


def assign(x, c):
    out = torch.empty(x.shape[0], dtype=torch.long)
    cf = c.float()
    n2 = (cf * cf).sum(1)
    for i in range(0, x.shape[0], 200_000):
        xx = x[i:i + 200_000]
        out[i:i + 200_000] = (n2[None, :] - 2 * xx @ cf.T).argmin(1)
    return out


def lists(c, delta, occ_cells):
    """mean list length over the occupied fine cells (exact double tests, slack delta)"""
    c = c.numpy()
    tot = 0
    lens = []
    for i in range(0, occ_cells.shape[0], 2048):
        cc = occ_cells[i:i + 2048]
        iz, iy, ix = cc // (G * G), (cc // G) % G, cc % G
        lo = np.stack([ix, iy, iz], 1) * cell - box
        hi = lo + cell
        a = c[None, :, :] - lo[:, None, :]
        b = hi[:, None, :] - c[None, :, :]
        far = np.maximum(np.abs(a), np.abs(b))
        near = np.maximum(np.maximum(-a, -b), 0)
        mx = (far ** 2).sum(2)
        mn = (near ** 2).sum(2)
        piv = mx.argmin(1)
        u = mx[np.arange(len(cc)), piv]
        s = 4 * delta * (np.sqrt(u) + delta)
        cp = c[piv]
        dlt = c[None, :, :] - cp[:, None, :]
        corner = np.where(dlt > 0, hi[:, None, :], lo[:, None, :])
        f = (-2 * dlt * corner).sum(2) + (c ** 2).sum(1)[None, :] - (cp ** 2).sum(1)[:, None]
        keep = (mn <= (u + s)[:, None]) & (f <= s[:, None])
        lens.append(keep.sum(1))
    return np.concatenate(lens)


shifts = []
cs = [c.clone()]
for it in range(ITERS):
    lab = assign(x, c)
    cnt = torch.bincount(lab, minlength=K).double()
    s = torch.zeros(K, 3, dtype=torch.float64).index_add_(0, lab, x.double())
    cn = torch.where(cnt[:, None] > 0, s / cnt[:, None].clamp_min(1), c)
    sh = (cn - c).norm(dim=1)
    shifts.append(float(sh.max()))
    print("iter %2d max shift %.5f (%.3f cells)  mean %.5f" % (it, shifts[-1], shifts[-1] / cell, float(sh.mean())))
    c = cn
    cs.append(c.clone())

# policy simulation: after the exchange of iteration i (centres cs[i+1]) decide skip/rebuild for iteration i+1
for alpha in (2.0, 3.0, 4.0, 6.0):
    for capf in (0.25, 0.5):
        ref, slack, builds, dl = cs[0], 0.0, 1, []
        for i in range(ITERS - 1):
            disp = float((cs[i + 1] - ref).norm(dim=1).max())
            if disp <= slack:
                continue
            builds += 1
            ref = cs[i + 1]
            d = alpha * shifts[i]
            slack = min(d, capf * cell) if shifts[i] <= capf * cell else 0.0
            dl.append(slack / cell)
        print("alpha %.1f cap %.2f: %d builds in %d iterations; slacks (cells): %s" %
              (alpha, capf, builds, ITERS, " ".join("%.3f" % t for t in dl)))

# list lengths against slack, weighted by rows
ix = ((x + box) / cell).long().clamp(0, G - 1)
cid = ix[:, 0] + G * ix[:, 1] + G * G * ix[:, 2]
occ, rows = torch.unique(cid, return_counts=True)
sel = torch.randperm(occ.shape[0])[:20000]
occ_s, rows_s = occ[sel].numpy(), rows[sel].numpy().astype(np.float64)
for df in (0.0, 0.02, 0.05, 0.1, 0.15, 0.25, 0.5):
    ln = lists(cs[-1], df * cell, occ_s)
    print("slack %.2f cells: mean list %.3f (row-weighted %.3f), max %d, >7: %.3f%%" %
          (df, ln.mean(), (ln * rows_s).sum() / rows_s.sum(), ln.max(), 100.0 * (ln > 7).mean()))
