"""Where KMeans.fit's wall time goes (10 M rotations, K = 1000, 20 fixed iterations, pinned host input).
usage: python scratch/e2e_anatomy.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-modal-regression_b200"), ROOT]
import torch
from bench import kmeans_chunks, synth_rotations, N_ROT, N_CHUNKS, K_DICT
from bdpose import ops, kmeans, _lib as L
dev = torch.device("cuda", 0)
xs = kmeans_chunks(range(N_CHUNKS), dev)
init = synth_rotations(N_ROT // N_CHUNKS, 100, dev, torch.float64)[:K_DICT].clone()
x_host = xs.cpu().pin_memory()
init_np = init.cpu().numpy()
del xs
steps = 20


def tick(acc, name, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t - t0) * 1e3
    return t


for rep in range(3):
    acc = {}
    torch.cuda.synchronize()
    t = time.perf_counter()
    x = x_host.to(dev, torch.float64, non_blocking=True)
    ini = torch.as_tensor(init_np, dtype=torch.float64).to(dev)
    t = tick(acc, "h2d", t)
    fs = kmeans.FitSetup(x, ini, kmeans.LOCAL, 1e-4, True)
    t = tick(acc, "fit_setup (4 passes + host reads)", t)
    state = kmeans.LloydState(x.shape[0], K_DICT, 3, dev)
    grid = ops.KeyGrid(fs.centers)
    t = tick(acc, "state + first grid", t)
    loop = kmeans.LloydLoop(fs.x, fs.centers, state.labels, fs.hb, grid, kmeans.LOCAL, fs.tol_abs, box=fs.max_abs)
    t = tick(acc, "loop setup (exchange, fixed geometry, occupancy)", t)
    loop.iterate(steps, check=False)
    t = tick(acc, "iterations", t)
    centers = loop.centers.clone()
    state.acc_stats.zero_(); state.inertia.zero_()
    kmeans.lloyd_step(fs.x, centers, state, fs.hb, update=False, grid=grid)
    inertia = float(state.inertia)
    t = tick(acc, "final E-step + inertia read", t)
    cc = (centers + fs.mean).cpu().numpy()
    stage = kmeans._pinned_labels(state.labels.numel())
    stage.copy_(state.labels, non_blocking=True)
    t = tick(acc, "d2h", t)
    if rep:
        print("rep %d: total %.2f ms | %s" % (rep, sum(acc.values()), " | ".join("%s %.2f" % kv for kv in acc.items())))
# the public call
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    km = kmeans.KMeans(n_clusters=K_DICT, init=init_np, n_init=1, max_iter=steps, fixed_iters=steps, device=dev)
    km.copy_labels = False
    km.fit(x_host)
    torch.cuda.synchronize()
    print("KMeans.fit %.2f ms" % ((time.perf_counter() - t0) * 1e3))
