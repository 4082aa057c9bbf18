"""Does the E+M kernel run faster on rotations sorted by grid cell?  (same kernel, same grid: only the
order of the rows changes: coalesced cell-record loads, broadcast candidate records, uniform list
lengths inside a warp).  usage: python scratch/time_sorted.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-modal-regression_b200"), ROOT]
import torch
from bench import kmeans_chunks, synth_rotations, N_ROT, N_CHUNKS, K_DICT
from bdpose import ops, kmeans, _lib as L
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else N_ROT
xs = kmeans_chunks(range(N_CHUNKS), dev)[:n].contiguous()
init = synth_rotations(N_ROT // N_CHUNKS, 100, dev, torch.float64)[:K_DICT].clone()
fs = kmeans.FitSetup(xs, init, group=kmeans.LOCAL)
G = 64
lo = -fs.max_abs * (1 + 1e-4)
cell = 2 * fs.max_abs * (1 + 1e-4) / G
ci = torch.floor((fs.x - lo) / cell).long().clamp_(0, G - 1)
orders = {"unsorted": None,
          "linear cell order": ci[:, 0] + G * (ci[:, 1] + G * ci[:, 2]),
          "coarse-major (4^3 blocks)": ((ci[:, 0] >> 2) + 16 * ((ci[:, 1] >> 2) + 16 * (ci[:, 2] >> 2))) * 64
                                       + (ci[:, 0] & 3) + 4 * ((ci[:, 1] & 3) + 4 * (ci[:, 2] & 3))}
steps = 20
hashes = []
for name, key in orders.items():
    x = fs.x if key is None else fs.x[torch.argsort(key)].contiguous()
    labels = torch.full((n,), -1, dtype=torch.int32, device=dev)
    loop = kmeans.LloydLoop(x, fs.centers, labels, fs.hb, ops.KeyGrid(fs.centers), kmeans.LOCAL, fs.tol_abs,
                            box=fs.max_abs)
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
    import hashlib
    h = hashlib.sha256(loop.c2[steps & 1].cpu().numpy().tobytes()).hexdigest()[:12]
    hashes.append(h)
    print("%-28s loop %.1f us/iter; E+M %s  mean %.1f  centres %s" % (
        name, e0.elapsed_time(e1) * 1e3 / steps, " ".join("%.0f" % v for v in em), sum(em) / steps, h))
print("same centres:", len(set(hashes)) == 1)
