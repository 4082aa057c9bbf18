"""Sharded key-grid build with a FAKE peer on the same GPU (world = 2, both grids / exchange buffers
local, the peer's flags pre-raised): separates the cost of the sharded code path from the cost of the
NVLink stores.  usage: python scratch/fake_shard.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-modal-regression_b200"), ROOT]
import torch
from bench import kmeans_chunks, synth_rotations, N_ROT, N_CHUNKS, K_DICT
from bdpose import ops, kmeans, _lib as L
dev = torch.device("cuda", 0)
lib = L.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else N_ROT // 2
xs = kmeans_chunks(range(N_CHUNKS), dev)[:n].contiguous()
init = synth_rotations(N_ROT // N_CHUNKS, 100, dev, torch.float64)[:K_DICT].clone()
fs = kmeans.FitSetup(xs, init, group=kmeans.LOCAL)
K, d = K_DICT, 3
A = K * (2 * d + 1) + 2
for world in (1, 2):
    labels = torch.full((n,), -1, dtype=torch.int32, device=dev)
    loop = kmeans.LloydLoop(fs.x, fs.centers, labels, fs.hb, ops.KeyGrid(fs.centers), kmeans.LOCAL, fs.tol_abs)
    nb = lib.bdp_kmeans_xchg_bytes(K, d) // 8
    xb = [torch.zeros(nb, dtype=torch.int64, device=dev) for _ in range(world)]
    gb = [torch.empty(loop.grid.nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    for g in gb:
        g.copy_(loop.grid.buf)           # prepared header + 0xFF cells
    if world == 2:
        xb[0][2 * A + 1] = 1 << 40       # flags[1]: the fake peer has "published" every iteration
        xb[0][2 * A + 8 + 1] = 1 << 40   # gflags[1]: and every slab
    steps = 20
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4 * steps)]
    for e in evs:
        e.record()
    torch.cuda.synchronize()
    ctl = torch.zeros(lib.bdp_kmeans_ctl_bytes(), dtype=torch.uint8, device=dev)
    c2 = torch.stack([fs.centers, fs.centers]).contiguous()
    xp = (C.c_void_p * world)(*[t.data_ptr() for t in xb])
    gp = (C.c_void_p * world)(*[t.data_ptr() for t in gb]) if world > 1 else None
    ev = (C.c_void_p * len(evs))(*[e.cuda_event for e in evs])
    st = lib.bdp_kmeans_run(fs.x.data_ptr(), n, d, c2.data_ptr(), K, gb[0].data_ptr(), loop.grid.nbytes, gp,
                            labels.data_ptr(), xp, None, world, 0, fs.hb, 0, steps, 0, 1, fs.tol_abs,
                            ctl.data_ptr(), ev, loop.cells.data_ptr(), loop.n_cells, L.stream_ptr())
    L.check(st, "run")
    torch.cuda.synchronize()
    b = [evs[4 * i].elapsed_time(evs[4 * i + 1]) * 1e3 for i in range(steps)]
    em = [evs[4 * i + 1].elapsed_time(evs[4 * i + 2]) * 1e3 for i in range(steps)]
    xf = [evs[4 * i + 2].elapsed_time(evs[4 * i + 3]) * 1e3 for i in range(steps)]
    print("world=%d cells=%d: build us %s" % (world, loop.n_cells, " ".join("%.0f" % v for v in b)))
    print("          E+M us %s" % " ".join("%.0f" % v for v in em))
    print("          xfin us %s" % " ".join("%.0f" % v for v in xf))
