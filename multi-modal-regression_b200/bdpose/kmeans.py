"""Pose-dictionary k-means on B200 (replaces sklearn.cluster.KMeans in learnKmeansDictionary.py:41-42
and the .predict of the pickled estimator used by binDeltaGenerators.py:27 / binDeltaLosses.py:35).

Semantics follow scikit-learn 1.9.0 Lloyd (`_kmeans_single_lloyd`, `lloyd_iter_chunked_dense`):
mean-centre X, E-step argmin ||x-c||^2 with lowest-index ties, M-step mean (sum * (1/count)),
empty clusters relocated to the points farthest from their centre, stop on unchanged labels or
sum ||dc||^2 <= tol * mean(var(X)), final E-step when not strictly converged, centres un-centred at
the end.

The iteration loop runs on the device: `bdp_kmeans_run` queues [key-grid build, E+M step,
exchange + finalise] for a batch of iterations; the stopping rules are evaluated by the finalise
kernel and a stopped fit turns the rest of the batch into no-ops, so the host reads one 40-byte
status per batch (none at all with `fixed_iters`).

Multi-GPU: rows are sharded over ranks, centres replicated.  The per-iteration exchange is fused
into the finalise kernel: every rank's int64 fixed-point accumulators [K, 2d+1] (+ the changed-label
counter) live in symmetric memory (torch.distributed._symmetric_memory: CUDA VMM handles mapped
into every process of the node) and each rank sums all peers' words straight over NVLink behind a
flag handshake — no NCCL call in the loop.  Integer sums are order-independent, so every rank (and
every world size) derives bit-identical centres.  Where symmetric memory cannot be set up the
exchange falls back to one NCCL all-reduce per iteration (`exchange="nccl"`).
"""
import os

import numpy as np
import torch

from . import _lib as L
from . import ops


def _fix_hi_bits(max_abs):
    """Largest fixed-point scale with |x| * 2^bits < 2^31."""
    e = 0 if max_abs <= 0 else int(np.floor(np.log2(max_abs))) + 1   # |x| < 2^e
    return int(max(0, min(30, 31 - max(e, 0) - 1)))


class LloydState:
    """Device buffers of one fit (allocated once, reused every iteration)."""

    def __init__(self, N, K, d, dev):
        self.labels = torch.full((N,), -1, dtype=torch.int32, device=dev)
        # [K*(2d+1)] accumulators followed by {changed, unused}: one all-reduce covers both
        self.acc_stats = torch.zeros(K * (2 * d + 1) + 2, dtype=torch.int64, device=dev)
        self.acc = self.acc_stats[: K * (2 * d + 1)]
        self.stats = self.acc_stats[K * (2 * d + 1):]
        self.inertia = torch.zeros(1, dtype=torch.float64, device=dev)
        self.shift2 = torch.zeros(1, dtype=torch.float64, device=dev)
        self.n_empty = torch.zeros(1, dtype=torch.int64, device=dev)


def lloyd_step(x, centers, state, fix_hi_bits, update=True, grid=None, want_inertia=False):
    """One E(+M-accumulate) step on this rank's shard.  Adds into state.acc / stats / inertia.
    grid: an ops.KeyGrid buffer to rebuild for `centers` and query through (candidate pruning), or
    None for the brute-force scan — the labels are the same either way."""
    N, d = x.shape
    # the inertia is only read after the final E-step (update=False); M-steps skip it
    inertia = None if (update and not want_inertia) else state.inertia
    with torch.cuda.device(x.device):
        if grid is None:
            st = L.lib().bdp_kmeans_lloyd_step(L.ptr(x), N, d, L.ptr(centers), centers.shape[0],
                                               L.ptr(state.labels), L.ptr(state.acc), fix_hi_bits,
                                               L.ptr(state.stats), L.ptr(inertia),
                                               1 if update else 0, L.stream_ptr())
        else:
            grid.rebuild(centers)
            st = L.lib().bdp_kmeans_lloyd_step_grid(L.ptr(x), N, d, L.ptr(grid.centers),
                                                    centers.shape[0], L.ptr(grid.buf), grid.nbytes,
                                                    L.ptr(state.labels), L.ptr(state.acc),
                                                    fix_hi_bits, L.ptr(state.stats),
                                                    L.ptr(inertia), 1 if update else 0,
                                                    L.stream_ptr())
    L.check(st, "bdp_kmeans_lloyd_step")


def lloyd_iteration(x, centers, centers_new, state, fix_hi_bits, grid=None, do_finalize=True):
    """zero accumulators + key-grid rebuild + E/M step (+ finalisation) in ONE library call."""
    N, d = x.shape
    K = centers.shape[0]
    lib = L.lib()
    with torch.cuda.device(x.device):
        st = lib.bdp_kmeans_iteration(
            x.data_ptr(), N, d, centers.data_ptr(), K,
            None if grid is None else grid.buf.data_ptr(), 0 if grid is None else grid.nbytes,
            state.labels.data_ptr(), state.acc_stats.data_ptr(), fix_hi_bits, None, 1,
            centers_new.data_ptr() if do_finalize else None, state.shift2.data_ptr(),
            state.n_empty.data_ptr(), L.stream_ptr())
    L.check(st, "bdp_kmeans_iteration")


def finalize(state, centers_old, centers_new, fix_hi_bits, shift2=None, n_empty=None):
    K, d = centers_old.shape
    shift2 = state.shift2 if shift2 is None else shift2
    n_empty = state.n_empty if n_empty is None else n_empty
    with torch.cuda.device(centers_old.device):
        st = L.lib().bdp_kmeans_finalize(L.ptr(state.acc), K, d, fix_hi_bits, L.ptr(centers_old),
                                         L.ptr(centers_new), L.ptr(shift2), L.ptr(n_empty),
                                         L.stream_ptr())
    L.check(st, "bdp_kmeans_finalize")


LOCAL = "local"     # group sentinel: this process alone, even when torch.distributed is initialised


def _dist_on(group):
    import torch.distributed as dist
    if isinstance(group, str) and group == LOCAL:
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def _relocate_empty(x, centers_old, state, fix_hi_bits, group):
    """sklearn _relocate_empty_clusters_dense on the accumulators (rare path).  The n_empty points
    farthest from their old centre each seed one empty cluster (in decreasing-distance order;
    sklearn's order inside the top-n set is whatever np.argpartition leaves, identical for
    n_empty == 1)."""
    import torch.distributed as dist
    K, d = centers_old.shape
    W = 2 * d + 1
    acc = state.acc.view(K, W)
    empty = torch.nonzero(acc[:, 2 * d] == 0).reshape(-1)
    n_empty = int(empty.numel())
    if n_empty == 0:
        return
    lab = state.labels.long()
    dist2 = ((x - centers_old[lab]) ** 2).sum(1)
    k = min(n_empty, dist2.numel())
    top_v, top_i = torch.topk(dist2, k)
    cand_x = x[top_i]
    cand_l = lab[top_i]
    if _dist_on(group):
        ws = dist.get_world_size(group)
        pad = n_empty - k
        pv = torch.cat([top_v, top_v.new_full((pad,), -1.0)])
        px = torch.cat([cand_x, cand_x.new_zeros((pad, d))])
        pl = torch.cat([cand_l, cand_l.new_zeros((pad,))])
        gv = [torch.empty_like(pv) for _ in range(ws)]
        gx = [torch.empty_like(px) for _ in range(ws)]
        gl = [torch.empty_like(pl) for _ in range(ws)]
        dist.all_gather(gv, pv, group=group)
        dist.all_gather(gx, px, group=group)
        dist.all_gather(gl, pl, group=group)
        allv, allx, alll = torch.cat(gv), torch.cat(gx), torch.cat(gl)
        top_v, sel = torch.topk(allv, n_empty)
        cand_x, cand_l = allx[sel], alll[sel]
    if float(top_v.max()) == 0.0:
        return   # more clusters than distinct samples: sklearn leaves things alone
    scale_hi = float(2.0 ** fix_hi_bits)
    for e in range(min(n_empty, cand_x.shape[0])):
        if float(top_v[e]) < 0:
            break
        xs = cand_x[e] * scale_hi
        hi = torch.floor(xs)
        lo = torch.trunc((xs - hi) * 4294967296.0)
        hi, lo = hi.to(torch.int64), lo.to(torch.int64)
        new_id, old_id = int(empty[e]), int(cand_l[e])
        acc[old_id, 0:2 * d:2] -= hi
        acc[old_id, 1:2 * d:2] -= lo
        acc[old_id, 2 * d] -= 1
        acc[new_id, 0:2 * d:2] = hi
        acc[new_id, 1:2 * d:2] = lo
        acc[new_id, 2 * d] = 1


class Exchange:
    """The exchange buffer of this rank ([2][K(2d+1)+2] int64 accumulators + flags, see
    include/bdpose.h) and this process's addresses of every peer's buffer."""
    _cache = {}

    def __init__(self, K, d, dev, group):
        import torch.distributed as dist
        lib = L.lib()
        self.K, self.d = K, d
        self.A = K * (2 * d + 1) + 2
        self.nbytes = lib.bdp_kmeans_xchg_bytes(K, d)
        if self.nbytes < 0:
            raise RuntimeError("kmeans: unsupported dictionary shape [%d, %d]" % (K, d))
        self.world = dist.get_world_size(group) if _dist_on(group) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.mode, self.why = "local", None
        self.mc = None
        n = self.nbytes // 8
        if self.world > 1:
            if self.world > L.KMEANS_MAX_RANKS:
                self.mode, self.why = "nccl", "more than %d ranks" % L.KMEANS_MAX_RANKS
            elif os.environ.get("BDPOSE_KMEANS_EXCHANGE", "p2p") == "nccl":
                self.mode, self.why = "nccl", "BDPOSE_KMEANS_EXCHANGE=nccl"
            else:
                try:
                    import torch.distributed._symmetric_memory as symm
                    pg = group if group is not None else dist.group.WORLD
                    self.buf = symm.empty(n, dtype=torch.int64, device=dev)
                    self.handle = symm.rendezvous(self.buf, pg)
                    ptrs = [int(p) for p in self.handle.buffer_ptrs]
                    if len(ptrs) != self.world or any(p == 0 for p in ptrs):
                        raise RuntimeError("rendezvous returned %r" % (ptrs,))
                    self.ptrs = ptrs
                    if os.environ.get("BDPOSE_KMEANS_NVLS", "0") == "1":
                        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
                        self.mc = mc if mc else None
                    self.mode = "p2p-nvls" if self.mc else "p2p"
                except Exception as e:      # no VMM / fabric / pidfd support on this box
                    self.mode, self.why = "nccl", "symmetric memory unavailable: %s" % (e,)
            # every rank must take the same path
            flag = torch.tensor([1 if self.mode == "nccl" else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
            if int(flag) and self.mode != "nccl":
                self.mode, self.why = "nccl", "a peer has no symmetric memory"
        if self.mode in ("local", "nccl"):
            self.buf = torch.zeros(n, dtype=torch.int64, device=dev)
            self.ptrs = [self.buf.data_ptr()]
        self.group = group
        self.reset()

    def reset(self):
        """Zero accumulators and flags; with peers: nobody may publish before everybody has zeroed."""
        import torch.distributed as dist
        self.buf.zero_()
        if self.world > 1:
            torch.cuda.synchronize(self.buf.device)
            dist.barrier(group=self.group)

    def acc(self, parity):
        return self.buf[parity * self.A:(parity + 1) * self.A]

    def shared_grid(self, nbytes):
        """A key-grid buffer every rank can store into (sharded build), or None when the exchange
        is not over peer memory.  Returns (uint8 tensor, ctypes array of every rank's address)."""
        import ctypes as C
        if not self.mode.startswith("p2p") or os.environ.get("BDPOSE_KMEANS_SHARD_GRID", "1") == "0":
            return None
        cur = getattr(self, "_grid", None)
        if cur is None or cur[0].numel() < nbytes:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            pg = self.group if self.group is not None else dist.group.WORLD
            buf = symm.empty(nbytes, dtype=torch.uint8, device=self.buf.device)
            hdl = symm.rendezvous(buf, pg)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            cur = self._grid = (buf, (C.c_void_p * self.world)(*ptrs), hdl)
        return cur[0], cur[1]

    def ptr_array(self):
        import ctypes as C
        if self.mode in ("local", "nccl"):
            return (C.c_void_p * 1)(self.ptrs[0]), 1, 0
        return (C.c_void_p * self.world)(*self.ptrs), self.world, self.rank

    @classmethod
    def get(cls, K, d, dev, group):
        key = (K, d, dev.index, group if isinstance(group, str) else (id(group) if group is not None else None))
        ex = cls._cache.get(key)
        if ex is None:
            ex = cls._cache[key] = cls(K, d, dev, group)
        else:
            ex.reset()
        return ex


class LloydLoop:
    """Lloyd iterations of one fit through bdp_kmeans_run (the device-side loop).

        loop = LloydLoop(x, centers, labels, hb, grid, group, tol_abs)
        loop.iterate(20)                 # fixed work: no stopping rules, no host synchronisation
        loop.iterate(300, check=True)    # scikit-learn's stopping rules, one status read per batch
        loop.centers, loop.n_iter, loop.strict, loop.mode

    x [N_local, d] fp64 (this rank's shard, already centred), centers [K, d] fp64, labels [N_local]
    int32 (previous labels in / new labels out)."""

    def __init__(self, x, centers, labels, hb, grid, group, tol_abs, box=None):
        lib = L.lib()
        self.box = box           # max |coordinate| over ALL ranks' rows if the caller knows it (FitSetup)
        self.x, self.labels, self.hb, self.grid, self.group = x, labels, hb, grid, group
        self.tol_abs = tol_abs
        self.dev = x.device
        self.N, self.d = x.shape
        self.K = centers.shape[0]
        self.ex = Exchange.get(self.K, self.d, self.dev, group)
        self.mode = self.ex.mode
        self.ctl = torch.zeros(lib.bdp_kmeans_ctl_bytes(), dtype=torch.uint8, device=self.dev)
        self.c2 = torch.stack([centers, centers]).contiguous()
        self.ptrs, self.world, self.rank = self.ex.ptr_array()
        # sharded key-grid build: the grid lives in peer-addressable memory, every rank builds one
        # slab of it per iteration and stores the slab into every rank's copy
        self.grid_ptrs = None
        if grid is not None:
            sg = self.ex.shared_grid(grid.nbytes)
            if sg is not None:
                self.grid = ops.KeyGrid(centers, buf=sg[0], build=False)
                self.grid_ptrs = sg[1]
        # fixed-geometry grid: laid over the bounding box of the ROWS once per fit, and only the coarse
        # cells that hold rows (on any rank) are rebuilt each iteration — the rest of the box is never
        # queried (uniform rotations fill 52 % of their bounding cube).  The loop owns this grid.
        self.cells, self.n_cells = None, 0
        # rows in cell order (bdp_cellsort): the loop then works on its own sorted copy of the rows and
        # its own label array; finish() scatters the labels back into the caller's array
        self.perm, self.user_labels = None, labels
        if self.grid is not None and self.mode != "nccl" and \
                os.environ.get("BDPOSE_KMEANS_FIXED_GRID", "1") != "0":
            if self.grid_ptrs is None:
                self.grid = ops.KeyGrid(centers, build=False)
            self._fix_geometry(sort=os.environ.get("BDPOSE_KMEANS_SORT", "1") != "0")
        self._tmp_shift = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self._tmp_empty = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.n_iter, self.strict, self.stopped = 0, False, False
        # incremental M-step (exact: integer sums): on whenever the loop runs through bdp_kmeans_run
        # with the key grid; the NCCL fallback all-reduces its accumulator in place and recomputes
        self.incremental = self.grid is not None and self.mode != "nccl" and \
            os.environ.get("BDPOSE_KMEANS_INCREMENTAL", "1") != "0"

    def _fix_geometry(self, sort=True):
        import torch.distributed as dist
        lib = L.lib()
        g, x, dev = self.grid, self.x, self.dev
        inf = float("inf")
        if self.box is not None:
            # the cube [-max|x|, max|x|]^d (FitSetup has the global maximum already: no extra pass)
            lo, hi = x.new_full((self.d,), -float(self.box)), x.new_full((self.d,), float(self.box))
        else:
            if self.N > 0:
                lo, hi = torch.aminmax(x, dim=0)
            else:
                lo, hi = x.new_full((self.d,), inf), x.new_full((self.d,), -inf)
            if self.world > 1:
                dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
            lo, hi = lo.contiguous(), hi.contiguous()
        occ = torch.zeros(lib.bdp_keygrid_coarse_cells(self.K, self.d), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            st = lib.bdp_keygrid_prepare(lo.data_ptr(), hi.data_ptr(), self.K, self.d, g.buf.data_ptr(),
                                         g.nbytes, L.stream_ptr())
            L.check(st, "bdp_keygrid_prepare")
            if sort and self.N > 0:
                nws = lib.bdp_cellsort_workspace_bytes(self.N, self.K, self.d)
                ws = torch.empty(nws, dtype=torch.uint8, device=dev)
                perm = torch.empty(self.N, dtype=torch.int32, device=dev)
                xs = torch.empty_like(x)
                st = lib.bdp_cellsort(x.data_ptr(), self.N, self.d, self.K, g.buf.data_ptr(), g.nbytes,
                                      occ.data_ptr(), ws.data_ptr(), nws, perm.data_ptr(), xs.data_ptr(),
                                      L.stream_ptr())
                L.check(st, "bdp_cellsort")
                self.x, self.perm = xs, perm
                self.labels = torch.full((self.N,), -1, dtype=torch.int32, device=dev)
                del ws
            else:
                st = lib.bdp_keygrid_occupancy(L.ptr(x), self.N, self.d, self.K, g.buf.data_ptr(), g.nbytes,
                                               occ.data_ptr(), L.stream_ptr())
                L.check(st, "bdp_keygrid_occupancy")
        if self.world > 1:
            # (also orders every rank's prepare before any peer's first slab store into its grid)
            dist.all_reduce(occ, op=dist.ReduceOp.MAX, group=self.group)
        self.cells = torch.nonzero(occ).reshape(-1).to(torch.int32).contiguous()
        self.n_cells = int(self.cells.numel())
        self.n_coarse = int(occ.numel())

    def finish(self):
        """Labels of the loop's rows back into the caller's array, in the caller's row order."""
        if self.perm is not None:
            with torch.cuda.device(self.dev):
                st = L.lib().bdp_scatter_i32(self.labels.data_ptr(), self.perm.data_ptr(), self.N,
                                             self.user_labels.data_ptr(), L.stream_ptr())
            L.check(st, "bdp_scatter_i32")
        return self.user_labels

    def reset(self, centers):
        """Start over from `centers` (same data): accumulators, flags, labels and status cleared."""
        self.c2[0].copy_(centers)
        self.c2[1].copy_(centers)
        self.ctl.zero_()
        self.labels.fill_(-1)
        self.ex.reset()
        self.n_iter, self.strict, self.stopped = 0, False, False

    @property
    def centers(self):
        return self.c2[self.n_iter & 1]

    def status(self):
        raw = bytes(self.ctl[:40].cpu().numpy())
        return L.KMeansStatus.from_buffer_copy(raw)

    def launch(self, i0, n, check, em_events=None):
        """Queue iterations i0 .. i0+n-1 on the current stream (no host synchronisation).
        em_events: 4n recorded-once torch.cuda.Event(enable_timing=True) objects, recorded per
        iteration before the grid build, before / after the E+M kernel and after the exchange kernel
        (benchmark instrumentation)."""
        import ctypes as C
        lib = L.lib()
        g = self.grid
        ev = None
        if em_events is not None:
            ev = (C.c_void_p * len(em_events))(*[e.cuda_event for e in em_events])
        with torch.cuda.device(self.dev):
            st = lib.bdp_kmeans_run(self.x.data_ptr(), self.N, self.d, self.c2.data_ptr(), self.K,
                                    None if g is None else g.buf.data_ptr(), 0 if g is None else g.nbytes,
                                    self.grid_ptrs, self.labels.data_ptr(), self.ptrs, self.ex.mc,
                                    self.world, self.rank,
                                    self.hb, i0, n, 1 if check else 0, 1 if self.incremental else 0,
                                    self.tol_abs, self.ctl.data_ptr(), ev,
                                    None if self.cells is None else self.cells.data_ptr(), self.n_cells,
                                    L.stream_ptr())
        L.check(st, "bdp_kmeans_run")

    def _launch_nccl(self, i0, check):
        # exchange fallback: the same kernels, the sums taken by one NCCL all-reduce per iteration
        import torch.distributed as dist
        lib = L.lib()
        ex, g, c2 = self.ex, self.grid, self.c2
        cur = i0 & 1
        acc = ex.acc(cur)
        with torch.cuda.device(self.dev):
            if self.N > 0:
                if g is not None:
                    g.rebuild(c2[cur])
                    st = lib.bdp_kmeans_lloyd_step_grid(
                        self.x.data_ptr(), self.N, self.d, c2[cur].data_ptr(), self.K, g.buf.data_ptr(),
                        g.nbytes, self.labels.data_ptr(), acc.data_ptr(), self.hb,
                        acc[ex.A - 2:].data_ptr(), None, 1, L.stream_ptr())
                else:
                    st = lib.bdp_kmeans_lloyd_step(
                        self.x.data_ptr(), self.N, self.d, c2[cur].data_ptr(), self.K,
                        self.labels.data_ptr(), acc.data_ptr(), self.hb, acc[ex.A - 2:].data_ptr(), None,
                        1, L.stream_ptr())
                L.check(st, "bdp_kmeans_lloyd_step")
            dist.all_reduce(acc, group=self.group)
            st = lib.bdp_kmeans_exchange_finalize(
                self.ptrs, None, 1, 0, self.K, self.d, self.hb, cur, i0 + 1, 1 if check else 0, 0,
                self.tol_abs, c2[cur].data_ptr(), c2[cur ^ 1].data_ptr(), self.ctl.data_ptr(),
                L.stream_ptr())
        L.check(st, "bdp_kmeans_exchange_finalize")

    def _host_relocation(self, st, check):
        """An empty cluster (rare): scikit-learn relocates it to the point farthest from its centre.
        Host-driven, on the globally summed accumulator of that iteration."""
        import torch.distributed as dist
        ex = self.ex
        g = int(st.iter_done) - 1
        p = g & 1
        acc = ex.acc(p)
        if ex.mode != "nccl":
            # this rank's sums persist (incremental M-step) and the peers may still read them: the
            # relocation works on a copy of the GLOBAL sums
            acc = acc.clone()
            if ex.world > 1:
                dist.all_reduce(acc, group=self.group)
        view = _AccView(acc[:ex.A - 2], self.labels)
        _relocate_empty(self.x, self.c2[p], view, self.hb, self.group)
        finalize(view, self.c2[p], self.c2[p ^ 1], self.hb, shift2=self._tmp_shift,
                 n_empty=self._tmp_empty)
        shift2 = float(self._tmp_shift)
        self.ctl.view(torch.int32)[0] = L.KMEANS_RUNNING
        self.n_iter = g + 1
        if ex.world > 1 and ex.mode != "nccl":
            # the summed accumulator must not be read as a partial sum by a late peer: all ranks leave
            # the host path together
            torch.cuda.synchronize(self.dev)
            dist.barrier(group=self.group)
        if check and int(st.changed) == 0:
            self.strict = self.stopped = True
        elif check and shift2 <= self.tol_abs:
            self.stopped = True

    def iterate(self, n, check=False, batch=8):
        """Up to n more iterations.  check=False: fixed work (the only early exit is an empty cluster,
        seen when the status is read after the batch).  Returns True when a stopping rule fired."""
        end = self.n_iter + n
        while self.n_iter < end and not self.stopped:
            m = end - self.n_iter
            if check:
                m = min(m, batch)
            if self.mode == "nccl":
                m = 1
                self._launch_nccl(self.n_iter, check)
            else:
                self.launch(self.n_iter, m, check)
            st = self.status()       # one small synchronising read per batch (not per iteration)
            if st.state == L.KMEANS_RUNNING:
                self.n_iter += m
            elif st.state == L.KMEANS_NEEDS_HOST:
                self._host_relocation(st, check)
            else:
                self.strict = st.state == L.KMEANS_STRICT
                self.n_iter = int(st.iter_done)
                self.stopped = True
        return self.stopped

class _AccView:
    """The (acc, labels) pair the relocation / finalise helpers take."""

    def __init__(self, acc, labels):
        self.acc, self.labels = acc, labels
        self.shift2 = self.n_empty = None


def _fit_stats(x, mode, out, mean=None, scale=1.0, y=None):
    N, d = x.shape
    with torch.cuda.device(x.device):
        st = L.lib().bdp_fit_stats(x.data_ptr(), N, d, mode, L.ptr(mean), float(scale), L.ptr(y),
                                   out.data_ptr(), L.stream_ptr())
    L.check(st, "bdp_fit_stats")


class FitSetup:
    """What scikit-learn's `fit` does before the first Lloyd iteration, for this rank's shard:
    global mean / variance (X -= X.mean(0); tol = mean(var(X)) * tol) summed in fixed point — integer
    sums do not depend on how the rows are split over blocks or ranks, so the centred data, and with
    them every label and centre, are bit-identical for any world size.  Four passes over the shard
    (bdp_fit_stats: max |x|, two-limb column sums, centring + maxima, variance sums)."""

    def __init__(self, x, init, group=None, tol=1e-4, center=True):
        import torch.distributed as dist
        x = x.double().contiguous()
        dev = x.device
        N, d = x.shape
        distributed = _dist_on(group)

        def allreduce(t, op=None):
            if distributed:
                dist.all_reduce(t, op=op or dist.ReduceOp.SUM, group=group)
            return t
        MAXOP = dist.ReduceOp.MAX if distributed else None
        use_kernels = x.is_cuda and 1 <= d <= 8

        def absmax(t):
            if use_kernels:
                o = torch.zeros(2, dtype=torch.int64, device=dev)     # bit pattern of a double >= 0
                _fit_stats(t, 0, o)
                return o[:1]
            return (t.abs().max().reshape(1) if t.numel() else t.new_zeros(1)).view(torch.int64)
        n_tot = allreduce(torch.tensor([N], dtype=torch.int64, device=dev)).double()
        # (non-negative doubles order like their bit patterns: the MAX all-reduce runs on int64)
        amax = float(allreduce(absmax(x), op=MAXOP).view(torch.float64))
        hb0 = _fix_hi_bits(amax)
        if use_kernels:
            limbs = torch.zeros(2 * d, dtype=torch.int64, device=dev)
            _fit_stats(x, 1, limbs, scale=2.0 ** hb0)
            limbs = limbs.view(2, d)
        else:
            xs = x * float(2.0 ** hb0)
            fl = torch.floor(xs)
            limbs = torch.stack([fl.to(torch.int64).sum(0),
                                 torch.trunc((xs - fl) * 4294967296.0).to(torch.int64).sum(0)])
            del xs, fl
        allreduce(limbs)
        mean = (limbs[0].double() * float(2.0 ** -hb0) + limbs[1].double() * float(2.0 ** -(hb0 + 32))) / n_tot
        if use_kernels:
            xc = torch.empty_like(x)
            mx = torch.zeros(2, dtype=torch.int64, device=dev)
            _fit_stats(x, 2, mx, mean=mean.contiguous(), y=xc)
            mx = allreduce(mx, op=MAXOP).view(torch.float64)
            d2max, cmax = float(mx[0]), float(mx[1])
        else:
            xc = x - mean
            d2 = xc ** 2
            d2max = float(allreduce((d2.max().reshape(1) if N else x.new_zeros(1)).view(torch.int64),
                                    op=MAXOP).view(torch.float64))
            cmax = float(allreduce((xc.abs().max().reshape(1) if N else x.new_zeros(1)).view(torch.int64),
                                   op=MAXOP).view(torch.float64))
        sh = 40 if d2max <= 0 else int(max(0, min(40, np.floor(61 - np.log2(d2max * float(n_tot) + 1.0)))))
        if use_kernels:
            q = torch.zeros(d, dtype=torch.int64, device=dev)
            _fit_stats(xc, 3, q, scale=2.0 ** sh)
        else:
            q = torch.round(xc ** 2 * float(2.0 ** sh)).to(torch.int64).sum(0)
        allreduce(q)
        var = q.double() * float(2.0 ** -sh) / n_tot
        self.tol_abs = float(var.mean()) * tol
        self.mean = mean
        self.center = center
        if center:
            x = xc
            centers = (init.double().to(dev) - mean).contiguous()
            max_abs = cmax
        else:
            centers = init.double().to(dev).contiguous().clone()
            max_abs = amax
        self.hb = _fix_hi_bits(max_abs)
        self.max_abs = max_abs
        self.x, self.centers = x, centers
        self.group, self.allreduce = group, allreduce
        self.distributed = distributed


def kmeans_lloyd(x, init, max_iter=300, tol=1e-4, group=None, fixed_iters=None, center=True,
                 use_grid="auto", _backend=None):
    """Lloyd k-means on this rank's shard `x` [N_local, d] fp64 (CUDA) from explicit centres.

    Returns dict(centers [K,d] fp64, labels [N_local] int32, inertia float, n_iter int, exchange).
    With fixed_iters=n the convergence tests are skipped and exactly n E+M iterations run (the
    benchmark's fixed-work mode); otherwise sklearn's stopping rules apply.
    """
    # _backend: (lloyd_step, finalize) callables standing in for the CUDA entry points — used by the
    # CPU/gloo tests of the host-driven loop (tests/test_dist_gloo.py); the product path never sets it.
    if _backend is None:
        ops._need_cuda(x, init)
        step_fn, finalize_fn = lloyd_step, finalize
    else:
        step_fn, finalize_fn = _backend
        use_grid = False
    fs = FitSetup(x, init, group, tol, center)
    x, centers, hb, tol_abs, mean = fs.x, fs.centers, fs.hb, fs.tol_abs, fs.mean
    allreduce = fs.allreduce
    dev = x.device
    N, d = x.shape
    K = init.shape[0]

    state = LloydState(N, K, d, dev)
    centers_new = torch.empty_like(centers)
    grid = None
    if use_grid is True or (use_grid == "auto" and ops.KeyGrid.supported(K, d, N)):
        grid = ops.KeyGrid(centers, build=False)      # (every user rebuilds it for its own centres)
    strict = False
    n_iter = 0
    iters = fixed_iters if fixed_iters is not None else max_iter
    exchange = "host-loop"
    loop = None
    if _backend is None:
        loop = LloydLoop(x, centers, state.labels, hb, grid, group, tol_abs, box=fs.max_abs)
        loop.iterate(iters, check=fixed_iters is None)
        centers, n_iter, strict, exchange = loop.centers.clone(), loop.n_iter, loop.strict, loop.mode
        iters = 0
        # (the loop may work on a cell-sorted copy of the rows with its own label array: the final
        # E-step runs on those, the labels are scattered back at the end)
        x, state.labels = loop.x, loop.labels
    for it in range(iters):
        # host-driven loop (stand-in backends only): one all-reduce + one status read per iteration
        state.acc_stats.zero_()
        step_fn(x, centers, state, hb, update=True, grid=grid)
        allreduce(state.acc_stats)
        finalize_fn(state, centers, centers_new, hb)
        if fixed_iters is None:
            host = torch.cat([state.stats[:1].double(), state.n_empty.double(), state.shift2]).tolist()
            changed, n_empty, shift2 = int(host[0]), int(host[1]), host[2]
            if n_empty > 0:
                _relocate_empty(x, centers, state, hb, group)
                finalize_fn(state, centers, centers_new, hb)
                shift2 = float(state.shift2)
        else:
            changed, shift2 = 1, float("inf")
        centers, centers_new = centers_new, centers
        n_iter = it + 1
        if changed == 0:
            strict = True
            break
        if shift2 <= tol_abs:
            break
    # final E-step (labels consistent with the returned centres) + inertia
    state.acc_stats.zero_()
    state.inertia.zero_()
    if not strict:
        step_fn(x, centers, state, hb, update=False, grid=grid)
        inertia = allreduce(state.inertia.clone())
    else:
        lab = state.labels.long()
        inertia = allreduce(((x - centers[lab]) ** 2).sum().reshape(1))
    if loop is not None:
        state.labels = loop.finish()
    if center:
        centers = centers + mean
    return dict(centers=centers, labels=state.labels, inertia=float(inertia), n_iter=n_iter,
                exchange=exchange)


def kmeans_plusplus(x, K, seed=0, n_local_trials=None):
    """k-means++ seeding on the device (sklearn _kmeans_plusplusalgorithm; the random stream is
    torch's, so seeds are not interchangeable with sklearn's — parity runs pass explicit `init`)."""
    ops._need_cuda(x)
    x = x.double()
    N, d = x.shape
    g = torch.Generator(device=x.device)
    g.manual_seed(seed)
    if n_local_trials is None:
        n_local_trials = 2 + int(np.log(K))
    centers = torch.empty((K, d), dtype=torch.float64, device=x.device)
    first = int(torch.randint(N, (1,), generator=g, device=x.device))
    centers[0] = x[first]
    closest = ((x - centers[0]) ** 2).sum(1)
    pot = closest.sum()
    for c in range(1, K):
        r = torch.rand(n_local_trials, generator=g, device=x.device, dtype=torch.float64) * pot
        cum = torch.cumsum(closest, 0)
        cand = torch.searchsorted(cum, r).clamp_(max=N - 1)
        dc = ((x[None, :, :] - x[cand][:, None, :]) ** 2).sum(2)     # [trials, N]
        dc = torch.minimum(dc, closest[None, :])
        pots = dc.sum(1)
        best = int(torch.argmin(pots))
        centers[c] = x[cand[best]]
        closest = dc[best]
        pot = pots[best]
    return centers


_pinned = [None]


def _pinned_labels(n):
    """ONE pinned staging buffer, grown on demand (a buffer per distinct n would pin host memory for
    good in a process that fits many different shards)."""
    t = _pinned[0]
    if t is None or t.numel() < n:
        t = _pinned[0] = torch.empty(max(n, 1), dtype=torch.int32).pin_memory()
    return t[:n]


class KMeans:
    """Drop-in for the pickled sklearn estimator the reference passes around
    (learnKmeansDictionary.py:41-47): exposes n_clusters, cluster_centers_ [K,d] float64 (numpy),
    labels_, inertia_, n_iter_, fit(X), predict(X).  Pickles hold numpy arrays only."""

    def __init__(self, n_clusters=8, init="k-means++", n_init=1, max_iter=300, tol=1e-4,
                 verbose=0, random_state=0, n_jobs=None, device=None, group=LOCAL, fixed_iters=None):
        self.n_clusters = n_clusters
        self.init = init
        self.n_init = n_init
        self.max_iter = max_iter
        self.tol = tol
        self.verbose = verbose
        self.random_state = random_state
        self.n_jobs = n_jobs          # accepted and ignored (old sklearn API used by the reference)
        self.device = device
        # extensions: `group` = a torch.distributed process group whose ranks each pass their own rows
        # to fit() (default: this process alone); `fixed_iters` = run exactly that many iterations
        self.group = group
        self.fixed_iters = fixed_iters
        self.copy_labels = True       # False: labels_ is a view of a reused pinned buffer (benchmarks)

    def _dev(self):
        return torch.device(self.device) if self.device is not None else torch.device(
            "cuda", torch.cuda.current_device())

    def fit(self, X, y=None):
        dev = self._dev()
        if isinstance(X, torch.Tensor):
            x = X.to(dev, torch.float64, non_blocking=True)
        else:
            x = torch.as_tensor(np.ascontiguousarray(X), dtype=torch.float64).to(dev)
        best = None
        n_init = 1 if not isinstance(self.init, str) else max(1, int(self.n_init))
        for trial in range(n_init):
            if isinstance(self.init, str):
                if self.init != "k-means++":
                    raise NameError("Unknown init passed")
                init = kmeans_plusplus(x, self.n_clusters, seed=int(self.random_state or 0) + trial)
            else:
                init = torch.as_tensor(np.asarray(self.init), dtype=torch.float64).to(dev)
            r = kmeans_lloyd(x, init, max_iter=self.max_iter, tol=self.tol, group=self.group,
                             fixed_iters=self.fixed_iters)
            if self.verbose:
                print("kmeans trial %d: inertia %.6f after %d iterations" % (trial, r["inertia"],
                                                                             r["n_iter"]))
            if best is None or r["inertia"] < best["inertia"]:
                best = r
        self.cluster_centers_ = best["centers"].cpu().numpy()
        # labels come back through a pinned staging buffer (kept per size: a 40 MB pageable D2H copy
        # costs as much as the whole fit)
        lab = best["labels"]
        stage = _pinned_labels(lab.numel())
        stage.copy_(lab, non_blocking=True)
        torch.cuda.current_stream(lab.device).synchronize()
        self.labels_ = stage.numpy().copy() if self.copy_labels else stage.numpy()
        self.inertia_ = best["inertia"]
        self.n_iter_ = best["n_iter"]
        return self

    def predict(self, X):
        dev = self._dev()
        X = np.ascontiguousarray(X)
        x = torch.as_tensor(X).to(dev)
        lab, _, _ = ops.assign_nearest(x, torch.as_tensor(self.cluster_centers_).to(dev),
                                       want_residual=False, label_dtype=torch.int32)
        return lab.cpu().numpy()

    def fit_predict(self, X, y=None):
        return self.fit(X).labels_
