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
