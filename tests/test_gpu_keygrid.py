"""Key grid (candidate pruning) against the brute-force scan and the oracle: the pruned query must
return bit-identical labels, residuals and distances, and the Lloyd step bit-identical fixed-point
accumulators, on every kind of dictionary — including the ones that push points onto the slow path
(outside the grid, overflowing cells, degenerate extents)."""
import numpy as np
import pytest
import torch

import bdpose_oracle as O

pytestmark = pytest.mark.gpu


def rand_rot(rng, n):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 0] < 0] *= -1
    ang = 2 * np.arccos(np.clip(q[:, 0], -1, 1))
    ax = q[:, 1:] / np.maximum(np.linalg.norm(q[:, 1:], axis=1, keepdims=True), 1e-300)
    return ax * ang[:, None], q


def _both(cuda, x, c, want_sq=True):
    from bdpose import ops
    xt, ct = torch.from_numpy(x).to(cuda), torch.from_numpy(c).to(cuda)
    g = ops.KeyGrid(ct)
    a = ops.assign_nearest(xt, ct, want_sqdist=want_sq, grid=g)
    b = ops.assign_nearest(xt, ct, want_sqdist=want_sq, grid=None)
    assert torch.equal(a[0], b[0]), "labels differ: %d rows" % int((a[0] != b[0]).sum())
    assert torch.equal(a[1], b[1])
    if want_sq:
        assert torch.equal(a[2], b[2])
    return a[0].cpu().numpy()


@pytest.mark.parametrize("N,K,d,dt", [(5000, 8, 3, np.float32), (40_000, 16, 3, np.float32),
                                      (100_000, 200, 3, np.float64), (100_000, 1000, 3, np.float32),
                                      (30_000, 4096, 3, np.float64), (50_000, 16, 4, np.float32),
                                      (100_000, 200, 4, np.float64), (60_000, 1000, 4, np.float32)])
def test_grid_equals_brute_force_and_oracle(cuda, N, K, d, dt):
    rng = np.random.default_rng(N + K + d)
    aa, q = rand_rot(rng, N + K)
    pts = aa if d == 3 else q
    x = pts[:N].astype(dt)
    c = pts[N:].astype(np.float64)
    lab = _both(cuda, x, c)
    n_or = min(N, 20_000)
    ob, _ = O.predict_residual(x[:n_or], c)
    assert np.array_equal(lab[:n_or], ob)


def test_grid_kmeans_like_dictionary(cuda):
    """Keys that are cluster means (strictly inside the data hull: many points fall in the grid
    margin or outside the grid)."""
    from bdpose import kmeans
    rng = np.random.default_rng(3)
    X = rand_rot(rng, 200_000)[0]
    r = kmeans.kmeans_lloyd(torch.from_numpy(X).to(cuda), torch.from_numpy(X[:300].copy()).to(cuda),
                            fixed_iters=5)
    c = r["centers"].cpu().numpy()
    _both(cuda, X.astype(np.float32), c)
    _both(cuda, X * 3.0, c)                      # most points far outside the grid: slow path
    _both(cuda, X * 1e-3, c)                     # all points in a handful of cells


def test_grid_degenerate_dictionaries(cuda):
    rng = np.random.default_rng(5)
    x = rand_rot(rng, 40_000)[0]
    same = np.tile(x[:1], (64, 1))                              # zero extent: grid disabled
    lab = _both(cuda, x, same)
    assert np.all(lab == 0)
    plane = rand_rot(rng, 500)[0]; plane[:, 2] = 0.25           # zero extent along one axis
    _both(cuda, x, plane)
    dup = rand_rot(rng, 100)[0]; dup[50:] = dup[:50]            # every key duplicated: low index wins
    lab = _both(cuda, x, dup)
    assert lab.max() < 50
    tight = x[:1] + rng.standard_normal((1500, 3)) * 1e-3       # 1500 keys in a tiny ball: overflow
    far = np.concatenate([tight, rand_rot(rng, 20)[0]])
    _both(cuda, x, far)
    xb = x.copy(); xb[7] = np.nan; xb[9, 1] = np.inf            # non-finite rows take the slow path
    from bdpose import ops
    xt, ct = torch.from_numpy(xb).to(cuda), torch.from_numpy(plane).to(cuda)
    a = ops.assign_nearest(xt, ct, grid=ops.KeyGrid(ct))[0]
    b = ops.assign_nearest(xt, ct, grid=None)[0]
    ok = torch.ones(len(xb), dtype=torch.bool, device=cuda); ok[7] = ok[9] = False
    assert torch.equal(a[ok], b[ok])


def test_grid_near_ties(cuda):
    """Points a few fp64 ulps off bisectors: the exact pass over the candidate list must agree with
    the exact pass over the whole dictionary."""
    rng = np.random.default_rng(11)
    c = rand_rot(rng, 300)[0]
    x = [rand_rot(rng, 40_000)[0]]
    for _ in range(2000):
        a, b = rng.integers(0, 300, 2)
        x.append((0.5 * (c[a] + c[b]) + rng.standard_normal(3) * 1e-12)[None])
    _both(cuda, np.concatenate(x), c)


@pytest.mark.parametrize("K,d", [(200, 3), (1000, 3), (200, 4)])
def test_grid_lloyd_step_bit_identical(cuda, K, d):
    from bdpose import kmeans, ops
    rng = np.random.default_rng(K + d)
    aa, q = rand_rot(rng, 300_000)
    X = torch.from_numpy(aa if d == 3 else q).to(cuda)
    c = X[:K].clone().contiguous()
    hb = kmeans._fix_hi_bits(float(X.abs().max()))
    outs = []
    for grid in (None, ops.KeyGrid(c)):
        st = kmeans.LloydState(X.shape[0], K, d, cuda)
        kmeans.lloyd_step(X, c, st, hb, update=True, grid=grid, want_inertia=True)
        outs.append((st.labels.clone(), st.acc_stats.clone(), float(st.inertia)))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    assert abs(outs[0][2] - outs[1][2]) <= 1e-9 * abs(outs[0][2])


def test_grid_kmeans_fit_identical(cuda):
    from bdpose import kmeans
    rng = np.random.default_rng(17)
    X = torch.from_numpy(rand_rot(rng, 150_000)[0]).to(cuda)
    init = X[:200].clone()
    a = kmeans.kmeans_lloyd(X, init, max_iter=15, use_grid=True)
    b = kmeans.kmeans_lloyd(X, init, max_iter=15, use_grid=False)
    assert a["n_iter"] == b["n_iter"]
    assert torch.equal(a["labels"], b["labels"])
    assert torch.equal(a["centers"], b["centers"])


@pytest.mark.parametrize("N,chunk", [(250_000, 65_536), (70_001, 1 << 20), (5, 2)])
def test_host_to_host_pipeline(cuda, N, chunk):
    """assign_labels_host (chunked three-stream pipeline from pinned host memory) returns exactly
    what the one-shot device call returns."""
    import binDeltaGenerators as G
    rng = np.random.default_rng(N)
    aa = rand_rot(rng, N + 500)[0]
    y = torch.from_numpy(aa[:N].astype(np.float32)).pin_memory()
    c = torch.from_numpy(aa[N:]).to(cuda)
    b_h, r_h = G.assign_labels_host(y, c, chunk_rows=chunk)
    b_d, r_d = G.assign_labels(y.to(cuda), c)
    assert b_h.dtype == torch.int64 and not b_h.is_cuda
    assert torch.equal(b_h, b_d.cpu()) and torch.equal(r_h, r_d.cpu())


def _clustered(rng, n, n_blobs=12, spread=0.15):
    """Pascal-like pose data: rotations concentrated around a few viewpoints, plus one far outlier
    that stretches the bounding box (most coarse cells of the fixed-geometry grid hold no row)."""
    cen = rand_rot(rng, n_blobs)[0] * 0.6
    x = cen[rng.integers(0, n_blobs, n)] + rng.standard_normal((n, 3)) * spread
    x[0] = [3.0, -3.0, 2.5]
    return x


@pytest.mark.parametrize("data", ["uniform", "clustered"])
def test_fixed_geometry_fit_identical(cuda, data, monkeypatch):
    """The k-means loop's fixed-geometry grid (box of the rows, only the occupied coarse cells rebuilt,
    counters / fp32 keys renewed by the exchange kernel) against the per-iteration dictionary-box grid
    and the brute-force scan: same iteration count, labels and centres bit for bit."""
    from bdpose import kmeans
    rng = np.random.default_rng(23)
    Xn = rand_rot(rng, 200_000)[0] if data == "uniform" else _clustered(rng, 200_000)
    X = torch.from_numpy(Xn).to(cuda)
    init = X[1:301].clone()
    fs = kmeans.FitSetup(X, init, group=kmeans.LOCAL)
    lab = torch.full((X.shape[0],), -1, dtype=torch.int32, device=cuda)
    from bdpose import ops
    loop = kmeans.LloydLoop(fs.x, fs.centers, lab, fs.hb, ops.KeyGrid(fs.centers), kmeans.LOCAL, fs.tol_abs)
    assert loop.cells is not None and 0 < loop.n_cells <= loop.n_coarse
    if data == "clustered":
        assert loop.n_cells < loop.n_coarse // 2, (loop.n_cells, loop.n_coarse)
    assert loop.perm is not None and torch.equal(loop.x, fs.x[loop.perm.long()])     # rows in cell order
    outs = []
    for fixed, use_grid, srt in (("1", True, "1"), ("0", True, "1"), ("1", False, "1"), ("1", True, "0")):
        monkeypatch.setenv("BDPOSE_KMEANS_FIXED_GRID", fixed)
        monkeypatch.setenv("BDPOSE_KMEANS_SORT", srt)
        for kw in (dict(max_iter=25), dict(fixed_iters=9)):
            outs.append(kmeans.kmeans_lloyd(X, init, use_grid=use_grid, **kw))
    for a, b in ((0, 2), (0, 4), (0, 6), (1, 3), (1, 5), (1, 7)):
        assert outs[a]["n_iter"] == outs[b]["n_iter"]
        assert torch.equal(outs[a]["labels"], outs[b]["labels"])
        assert torch.equal(outs[a]["centers"], outs[b]["centers"])
        # (the inertia is a floating-point sum whose order differs between the kernels)
        assert outs[a]["inertia"] == pytest.approx(outs[b]["inertia"], rel=1e-12)


def test_prepared_grid_scans_until_built(cuda):
    """bdp_keygrid_prepare marks every cell "scan the dictionary": a query through a prepared but
    unbuilt grid is exact (the occupied-cell list is an optimisation, not a correctness condition),
    and bdp_keygrid_occupancy marks the coarse parents of exactly the cells the rows fall into."""
    from bdpose import ops, _lib as L
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rand_rot(rng, 20_000)[0]).to(cuda)
    c = x[:100].clone().contiguous()
    g = ops.KeyGrid(c, build=False)
    lo, hi = torch.aminmax(x, dim=0)
    lib = L.lib()
    L.check(lib.bdp_keygrid_prepare(lo.contiguous().data_ptr(), hi.contiguous().data_ptr(), 100, 3,
                                    g.buf.data_ptr(), g.nbytes, L.stream_ptr()), "prepare")
    a = ops.assign_nearest(x, c, grid=g)
    b = ops.assign_nearest(x, c, grid=None)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    st = ops.keygrid_stats(g, x)
    assert st["outside_grid"] == 0 and st["overflow_cells"] == x.shape[0]
    nc = lib.bdp_keygrid_coarse_cells(100, 3)
    occ = torch.zeros(nc, dtype=torch.int32, device=cuda)
    L.check(lib.bdp_keygrid_occupancy(x.data_ptr(), x.shape[0], 3, 100, g.buf.data_ptr(), g.nbytes,
                                      occ.data_ptr(), L.stream_ptr()), "occupancy")
    # the same map in numpy (fp32 FMA is not available there: compare with a tolerance-free superset
    # check instead — every marked cell holds a row within one fine cell of it, every row's cell is marked)
    hdr = np.frombuffer(bytes(g.buf[:160].cpu().numpy()), dtype=np.float64, count=12)
    org, inv = hdr[0:3], hdr[8:11]
    G = int(np.frombuffer(bytes(g.buf[96:100].cpu().numpy()), dtype=np.int32)[0])
    t = (x.cpu().numpy() - org) * inv
    fine = np.floor(t).astype(np.int64)
    assert fine.min() >= 0 and fine.max() < G
    frac = t - fine
    safe = ((frac > 1e-3) & (frac < 1 - 1e-3)).all(1)          # rows whose cell no rounding can change
    Gc = G // 4
    coarse = (fine[:, 0] // 4) + Gc * ((fine[:, 1] // 4) + Gc * (fine[:, 2] // 4))
    occ_h = occ.cpu().numpy()
    assert occ_h[coarse[safe]].all()
    allowed = set()
    for sx in (-2e-3, 2e-3):
        for sy in (-2e-3, 2e-3):
            for sz in (-2e-3, 2e-3):
                f = np.clip(np.floor(t + np.array([sx, sy, sz])).astype(np.int64), 0, G - 1)
                allowed |= set(((f[:, 0] // 4) + Gc * ((f[:, 1] // 4) + Gc * (f[:, 2] // 4))).tolist())
    assert set(np.nonzero(occ_h)[0].tolist()) <= allowed


def test_cellsort_is_a_permutation_in_cell_order(cuda):
    """bdp_cellsort: perm is a permutation, x_sorted = x[perm], the cell ids are non-decreasing along
    it, the occupied-cell marks equal bdp_keygrid_occupancy's, and bdp_scatter_i32 inverts the order."""
    from bdpose import ops, _lib as L
    rng = np.random.default_rng(9)
    x = torch.from_numpy(rand_rot(rng, 123_457)[0]).to(cuda)
    K, d, N = 300, 3, x.shape[0]
    g = ops.KeyGrid(x[:K].clone().contiguous(), build=False)
    lo, hi = torch.aminmax(x, dim=0)
    lib = L.lib()
    L.check(lib.bdp_keygrid_prepare(lo.contiguous().data_ptr(), hi.contiguous().data_ptr(), K, d,
                                    g.buf.data_ptr(), g.nbytes, L.stream_ptr()), "prepare")
    nc = lib.bdp_keygrid_coarse_cells(K, d)
    occ_a = torch.zeros(nc, dtype=torch.int32, device=cuda)
    occ_b = torch.zeros(nc, dtype=torch.int32, device=cuda)
    nws = lib.bdp_cellsort_workspace_bytes(N, K, d)
    ws = torch.empty(nws, dtype=torch.uint8, device=cuda)
    perm = torch.empty(N, dtype=torch.int32, device=cuda)
    xs = torch.empty_like(x)
    L.check(lib.bdp_cellsort(x.data_ptr(), N, d, K, g.buf.data_ptr(), g.nbytes, occ_a.data_ptr(),
                             ws.data_ptr(), nws, perm.data_ptr(), xs.data_ptr(), L.stream_ptr()), "cellsort")
    L.check(lib.bdp_keygrid_occupancy(x.data_ptr(), N, d, K, g.buf.data_ptr(), g.nbytes, occ_b.data_ptr(),
                                      L.stream_ptr()), "occupancy")
    p = perm.long()
    assert torch.equal(torch.sort(p).values, torch.arange(N, device=cuda))
    assert torch.equal(xs, x[p])
    assert torch.equal(occ_a, occ_b)
    hdr = np.frombuffer(bytes(g.buf[:160].cpu().numpy()), dtype=np.float64, count=12)
    G = int(np.frombuffer(bytes(g.buf[96:100].cpu().numpy()), dtype=np.int32)[0])
    t = (xs.cpu().numpy() - hdr[0:3]) * hdr[8:11]
    fine = np.floor(t).astype(np.int64)
    frac = t - fine
    safe = ((frac > 1e-3) & (frac < 1 - 1e-3)).all(1)          # rows whose cell no rounding can change
    cid = fine[:, 0] + G * (fine[:, 1] + G * fine[:, 2])
    assert (np.diff(cid[safe]) >= 0).all()
    lab = torch.arange(N, dtype=torch.int32, device=cuda)
    out = torch.full((N,), -7, dtype=torch.int32, device=cuda)
    L.check(lib.bdp_scatter_i32(lab.data_ptr(), perm.data_ptr(), N, out.data_ptr(), L.stream_ptr()), "scatter")
    assert torch.equal(out[p], lab)


def test_quaternion_fit_through_the_sorted_loop(cuda):
    """d = 4 (quaternion dictionaries, learnKmeansDictionary with quaternion targets): the loop's d = 4
    instantiations (fixed-geometry grid, occupied cells, cell sort, label scatter) against the
    brute-force fit — same iterations, labels and centres bit for bit; N not a multiple of the tile."""
    from bdpose import kmeans
    rng = np.random.default_rng(41)
    X = torch.from_numpy(rand_rot(rng, 90_007)[1]).to(cuda)
    init = X[:150].clone()
    a = kmeans.kmeans_lloyd(X, init, max_iter=12, use_grid=True)
    b = kmeans.kmeans_lloyd(X, init, max_iter=12, use_grid=False)
    assert a["n_iter"] == b["n_iter"]
    assert torch.equal(a["labels"], b["labels"])
    assert torch.equal(a["centers"], b["centers"])
