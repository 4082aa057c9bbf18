"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden vectors.
Tolerances (BASELINE.json north_star): bin indices / argmin labels / k-means labels bit-exact;
losses, deltas, errors and gradients 1e-5 relative in fp32 (small absolute floors noted inline)."""
import os
import pickle
import warnings

import numpy as np
import pytest
import torch

import bdpose_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def rand_rot(rng, n):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 0] < 0] *= -1
    ang = 2 * np.arccos(np.clip(q[:, 0], -1, 1))
    ax = q[:, 1:] / np.maximum(np.linalg.norm(q[:, 1:], axis=1, keepdims=True), 1e-300)
    return ax * ang[:, None], q


def close(a, b, rtol=RTOL, atol=0.0, msg=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=msg)


# ---------------------------------------------------------------------------------------------------
# (b) fused loss
# ---------------------------------------------------------------------------------------------------
def _fused(cuda, score, bins, pred, target, keys, mode, use_keys):
    from bdpose import ops
    s = torch.as_tensor(score).to(cuda).requires_grad_(True)
    p = torch.as_tensor(pred).to(cuda).requires_grad_(True)
    out = ops.bd_loss(s, torch.as_tensor(bins).to(cuda), p, torch.as_tensor(target).to(cuda),
                      None if keys is None else torch.as_tensor(keys).to(cuda), mode, use_keys)
    return s, p, out


EPS32 = 2.0 ** -24


def _rows64(mode, pred, target, keys, ind):
    """fp64 per-row pose loss of the oracle on the (fp32-composed) prediction + its conditioning
    1/(1 - c^2): the reference evaluates acos(c) and 1/sqrt(1 - c^2) in fp32, so ITS result carries
    a relative rounding noise of ~eps32/(1 - c^2) that no other implementation can reproduce."""
    p = torch.as_tensor(pred)
    if keys is not None and mode != "riem":
        p = (torch.as_tensor(keys).float()[ind] + p.float())
    p = p.double().requires_grad_(True)
    t = torch.as_tensor(target).double()
    if mode == "aa":
        rows = O.geodesic_loss_aa(p, t, reduce=False)
        c = torch.cos(rows.detach() / 2)
    elif mode == "quat":
        rows = O.geodesic_loss_quat(p, t, reduce=False)
        c = torch.cos(rows.detach() / 2)
    elif mode == "riem":
        K = torch.as_tensor(keys).double().reshape(-1, 3, 3)[ind]
        ang = p.norm(dim=1)
        ax = torch.nn.functional.normalize(p)
        A = (ax @ torch.from_numpy(O._PROJ).double()).view(-1, 3, 3)
        E = torch.eye(3, dtype=torch.float64) + torch.sin(ang)[:, None, None] * A + \
            (1 - torch.cos(ang))[:, None, None] * (A @ A)
        Rh = K @ E
        tr = (Rh * t.reshape(-1, 3, 3)).sum((1, 2))
        rows = torch.acos(torch.clamp((tr - 1) / 2, -1 + 1e-6, 1 - 1e-6))
        c = torch.cos(rows.detach())
    else:
        raise NameError(mode)
    rows.mean().backward()
    cond = 1.0 / (1.0 - c ** 2).clamp_min(1e-12)
    return rows.detach().numpy(), p.grad.numpy(), cond.numpy()


def _check_pose_grad(got, g64, g32, cond, msg):
    """got: ours (fp32).  Against the fp64 evaluation of the reference formula: 1e-5 of the row's
    gradient scale.  Against the reference's own fp32 result: the same plus its rounding noise."""
    got = got.detach().cpu().numpy().astype(np.float64)
    scale = np.abs(g64).max(axis=1, keepdims=True) + 1e-30
    assert np.all(np.abs(got - g64) <= 1e-5 * scale + 1e-12), msg + " vs fp64 oracle"
    if g32 is not None:
        tol = scale * (1e-5 + 16 * EPS32 * cond[:, None]) + 1e-12
        assert np.all(np.abs(got - g32) <= tol), msg + " vs reference fp32"


def test_losses_golden(cuda, golden):
    from bdpose import _lib as L, ops
    g = golden("losses")
    alpha = float(g["alpha"])
    qcenters = O.convert_dictionary(g["centers"]).astype(np.float32)
    ind = g["score"].argmax(1)
    cases = [
        ("simple", "res3", g["res_true"], None, L.POSE_MSE, False, None),
        ("geod_mse", "res3", g["ytrue_aa"], g["centers"].astype(np.float32), L.POSE_MSE, True, None),
        ("geod_aa", "res3", g["ytrue_aa"], g["centers"].astype(np.float32), L.POSE_GEODESIC_AA, True, "aa"),
        ("geod_q", "res4", g["ytrue_q"], qcenters, L.POSE_GEODESIC_Q, True, "quat"),
        ("riem", "res3", g["R_true"].reshape(-1, 9), g["key_rot"].reshape(-1, 9).astype(np.float32),
         L.POSE_RIEMANNIAN, True, "riem"),
    ]
    for name, rk, target, keys, mode, use_keys, kind in cases:
        s, p, out = _fused(cuda, g["score"], g["bin_true"], g[rk], target, keys, mode, use_keys)
        loss = out[0] + alpha * out[1]
        loss.backward()
        close(loss, g[name + "_loss"], msg=name)                 # 1e-5 relative
        # d(CE)/d(score): plain fp32 softmax arithmetic, 1e-5 of the gradient scale
        close(s.grad, g[name + "_g0"], atol=1e-5 * np.abs(g[name + "_g0"]).max(), msg=name)
        if kind is None:
            close(p.grad, g[name + "_g1"], atol=1e-5 * np.abs(g[name + "_g1"]).max(), msg=name)
        else:
            _, g64, cond = _rows64(kind, g[rk], target, keys, ind)
            _check_pose_grad(p.grad, alpha * g64, g[name + "_g1"], cond, name)
        # argmax with the planted tie (row 3: columns 5 and 9) picks the lowest index
        assert int(out[2][3]) == 5
    # stand-alone pose losses, mean and per-sample
    for key, pk, tk, mode, kind in (("geo_aa", "p_aa", "ytrue_aa", L.POSE_GEODESIC_AA, "aa"),
                                    ("geo_q", "p_q", "ytrue_q", L.POSE_GEODESIC_Q, "quat")):
        p = torch.from_numpy(g[pk]).to(cuda).requires_grad_(True)
        t = torch.from_numpy(g[tk]).to(cuda)
        v = ops.pose_loss(p, t, mode)
        v.backward()
        close(v, g[key + "_loss"], msg=key)
        rows64, g64, cond = _rows64(kind, g[pk], g[tk], None, None)
        _check_pose_grad(p.grad, g64, g[key + "_g0"], cond, key)
        rows = ops.pose_loss(p.detach(), t, mode, reduce=False)
        close(rows, rows64, rtol=1e-6, msg=key)                  # vs fp64 evaluation
        # vs the reference's fp32 rows: its acos noise is ~2 eps32 / sqrt(1 - c^2) absolute
        assert np.all(np.abs(rows.cpu().numpy() - g[key + "_rows"]) <=
                      1e-5 * rows64 + 8 * EPS32 * np.sqrt(cond)), key


@pytest.mark.parametrize("B,K", [(1, 200), (32, 200), (96, 200), (333, 24), (257, 203), (64, 1000),
                                 (50, 16), (5000, 200)])
def test_loss_vs_oracle_shapes(cuda, B, K):
    from bdpose import _lib as L
    rng = np.random.default_rng(B * 1000 + K)
    torch.manual_seed(B + K)
    score = (torch.randn(B, K) * 3).numpy()
    bins = rng.integers(0, K, B)
    centers = rand_rot(rng, K)[0].astype(np.float32)
    target = rand_rot(rng, B)[0].astype(np.float32)
    delta = (rng.standard_normal((B, 3)) * 0.3).astype(np.float32)
    s0 = torch.from_numpy(score).requires_grad_(True)
    d0 = torch.from_numpy(delta).requires_grad_(True)
    l1, l2 = O.bin_delta_terms(s0, d0, torch.from_numpy(bins), torch.from_numpy(target),
                               torch.from_numpy(centers), "aa")
    (l1 + 0.5 * l2).backward()
    s, p, out = _fused(cuda, score, bins, delta, target, centers, L.POSE_GEODESIC_AA, True)
    (out[0] + 0.5 * out[1]).backward()
    close(out[0], l1)
    close(out[1], l2)
    assert np.array_equal(out[2].cpu().numpy(), score.argmax(1))
    close(s.grad, s0.grad, atol=1e-5 * float(s0.grad.abs().max()))
    _, g64, cond = _rows64("aa", delta, target, centers, score.argmax(1))
    _check_pose_grad(p.grad, 0.5 * g64, d0.grad.numpy(), cond, "shapes")


def test_loss_two_forwards_one_backward(cuda):
    """learnGeodesicBDModel.py:116-120,183: two forwards, concatenated, then one backward."""
    from bdpose import _lib as L, ops
    rng = np.random.default_rng(5)
    K = 200
    centers = torch.from_numpy(rand_rot(rng, K)[0].astype(np.float32))
    sa, sb = torch.randn(48, K), torch.randn(48, K)
    da, db = torch.randn(48, 3) * 0.1, torch.randn(48, 3) * 0.1
    bins = torch.randint(0, K, (96,))
    tgt = torch.from_numpy(rand_rot(rng, 96)[0].astype(np.float32))

    def run(dev, fused):
        A = [t.clone().to(dev).requires_grad_(True) for t in (sa, sb, da, db)]
        score = torch.cat([A[0], A[1]])
        delta = torch.cat([A[2], A[3]])
        if fused:
            lc, lr, _ = ops.bd_loss(score, bins.to(dev), delta, tgt.to(dev), centers.to(dev),
                                    L.POSE_GEODESIC_AA, True)
        else:
            lc, lr = O.bin_delta_terms(score, delta, bins, tgt, centers, "aa")
        s = float(lr.detach().log())                     # python-float weighting of the script
        (lc + np.exp(-s) * lr + s).backward()
        return [a.grad.cpu() for a in A]
    for a, b in zip(run(cuda, True), run("cpu", False)):
        close(a, b, rtol=1e-4, atol=1e-4 * float(b.abs().max()))   # fp32 reference noise, see _rows64


def test_loss_module_api(cuda, golden, tmp_path):
    """The mirror classes with the reference's constructor signatures (binDeltaLosses.py)."""
    import binDeltaLosses as BL
    import axisAngle
    import quaternion
    g = golden("losses")
    kfile = str(tmp_path / "km.pkl")
    with open(kfile, "wb") as f:
        pickle.dump(_Dict(g["centers"]), f)
    alpha = float(g["alpha"])
    T = lambda k: torch.from_numpy(g[k]).to(cuda)
    crits = [
        ("simple", BL.SimpleLoss(alpha), "res3", T("res_true")),
        ("simple", BL.loss_m0(alpha), "res3", T("res_true")),
        ("geod_mse", BL.GeodesicLoss(alpha, kfile), "res3", T("ytrue_aa")),
        ("geod_aa", BL.GeodesicLoss(alpha, kfile, axisAngle.geodesic_loss()), "res3", T("ytrue_aa")),
        ("geod_aa", BL.loss_m1(alpha, kfile, axisAngle.geodesic_loss()), "res3", T("ytrue_aa")),
        ("geod_q", BL.GeodesicLossQ(alpha, kfile, quaternion.geodesic_loss()), "res4", T("ytrue_q")),
        ("riem", BL.RiemannianLoss(alpha, g["key_rot"]), "res3", T("R_true")),
    ]
    for name, crit, rk, target in crits:
        crit = crit.cuda()
        s = T("score").requires_grad_(True)
        r = T(rk).requires_grad_(True)
        loss = crit([s, r], [T("bin_true"), target])
        assert loss.dim() == 0
        loss.backward()
        close(loss, g[name + "_loss"], msg=name)
        close(r.grad, g[name + "_g1"], rtol=1e-4, atol=1e-4 * np.abs(g[name + "_g1"]).max(), msg=name)


class _Dict:
    def __init__(self, c):
        self.cluster_centers_ = c
        self.n_clusters = c.shape[0]


# ---------------------------------------------------------------------------------------------------
# (c) assignment
# ---------------------------------------------------------------------------------------------------
def test_label_generation_golden(cuda, golden):
    from bdpose import ops
    import binDeltaGenerators as G
    import quaternion
    g = golden("label_generation")
    c = g["centers"]
    ya = g["ydata_aa"].reshape(-1, 3)
    b, r = G.assign_labels(ya, c)
    assert b.dtype == torch.int64 and r.dtype == torch.float32
    assert np.array_equal(b.cpu().numpy(), g["gbd_bin"].reshape(-1))
    assert np.array_equal(r.cpu().numpy(), g["gbd_res"].reshape(-1, 3))
    qc = quaternion.convert_dictionary(c)
    close(qc, O.convert_dictionary(c), rtol=0, atol=1e-15)
    yq = g["ydata_q"].reshape(-1, 4)
    b, r = G.assign_labels(yq, qc)
    assert np.array_equal(b.cpu().numpy(), g["gbdq_bin"].reshape(-1))
    assert np.array_equal(r.cpu().numpy(), g["gbdq_res"].reshape(-1, 4))
    p, r = G.assign_soft_labels(yq, qc)
    close(p, g["xpbdq_bin"].reshape(-1, c.shape[0]), rtol=1e-5, atol=1e-30)
    close(r, g["xpbdq_res"].reshape(-1, 4), rtol=1e-5, atol=1e-7)
    b, r, rot = G.assign_labels_riemannian(ya, c)
    assert np.array_equal(b.cpu().numpy(), g["rbd_bin"].reshape(-1))
    # fp32 inputs: the reference rounds sin/cos/V.V to float32 inside get_R (numpy scalar rules);
    # we evaluate in fp64 -> agreement to float32 rounding (1e-6 abs on O(1) matrix entries)
    close(rot, g["rbd_rot"].reshape(-1, 3, 3), rtol=1e-5, atol=1e-6)
    close(r, g["rbd_res"].reshape(-1, 3), rtol=1e-5, atol=2e-6)
    b, r = ops.assign_quatdot(torch.from_numpy(g["qq"]).to(cuda), torch.from_numpy(g["qkeys"]).to(cuda))
    assert np.array_equal(b.cpu().numpy(), g["qbin"])
    assert np.array_equal(r.cpu().numpy(), g["qres"])


def test_riemannian_residual_fp64(cuda):
    """fp64 inputs: the kernel restates get_R / get_y exactly -> tight agreement, edge cases included."""
    from bdpose import ops
    rng = np.random.default_rng(3)
    y = rand_rot(rng, 500)[0]
    c = rand_rot(rng, 30)[0]
    y[0] = c[7]                       # identical to a key: residual exactly 0 (axis norm <= eps)
    y[1] = 0.0                        # identity rotation
    b, res, rot = O.riemannian_targets(y, c)
    key_rot = torch.from_numpy(np.stack([O.get_R(v) for v in c])).to(cuda)
    rot_g, res_g = ops.riemannian_residual(torch.from_numpy(y).to(cuda), key_rot,
                                           torch.from_numpy(b).to(cuda))
    close(rot_g, rot, rtol=0, atol=1e-7)      # outputs are float32
    close(res_g, res, rtol=1e-6, atol=1e-7)
    assert np.all(res_g[0].cpu().numpy() == 0)


@pytest.mark.parametrize("N,K,d,dt", [(1, 1, 3, np.float32), (1023, 7, 3, np.float32),
                                      (1025, 200, 3, np.float64), (5000, 1000, 3, np.float32),
                                      (3000, 2500, 3, np.float64), (2048, 64, 4, np.float32),
                                      (777, 200, 4, np.float64)])
def test_assign_vs_oracle(cuda, N, K, d, dt):
    from bdpose import ops
    rng = np.random.default_rng(N + K)
    aa, q = rand_rot(rng, N + K)
    pts = (aa if d == 3 else q)
    x = pts[:N].astype(dt)
    c = pts[N:].astype(np.float64)
    lab, res, sq = ops.assign_nearest(torch.from_numpy(x).to(cuda), torch.from_numpy(c).to(cuda),
                                      want_sqdist=True)
    ob, ores = O.predict_residual(x, c)
    assert np.array_equal(lab.cpu().numpy(), ob)
    assert np.array_equal(res.cpu().numpy(), ores)
    d2 = ((x.astype(np.float64) - c[ob]) ** 2).sum(1)
    close(sq, d2, rtol=1e-12, atol=1e-300)
    lab32, _, _ = ops.assign_nearest(torch.from_numpy(x).to(cuda), torch.from_numpy(c).to(cuda),
                                     want_residual=False, label_dtype=torch.int32)
    assert lab32.dtype == torch.int32 and np.array_equal(lab32.cpu().numpy(), ob)


def test_assign_ties_and_near_ties(cuda):
    """Exact ties (duplicated keys, points on a bisector) resolve to the lowest index; points a few
    fp64 ulps off a bisector — far inside the fp32 screening noise — still get the fp64 answer."""
    from bdpose import ops
    rng = np.random.default_rng(11)
    c = rand_rot(rng, 50)[0]
    c[31] = c[4]                                   # duplicate key: index 4 must win over 31
    x = np.concatenate([c, rand_rot(rng, 2000)[0]])
    mid = 0.5 * (c[10] + c[20])                    # exact bisector point in fp64 (if representable)
    eps = np.array([1e-13, 0, 0])
    x = np.concatenate([x, mid[None], (mid + eps)[None], (mid - eps)[None]])
    for j in range(200):                           # a cloud hugging bisectors at fp64 resolution
        a, b = rng.integers(0, 50, 2)
        x = np.concatenate([x, (0.5 * (c[a] + c[b]) + rng.standard_normal(3) * 1e-12)[None]])
    lab, _, _ = ops.assign_nearest(torch.from_numpy(x).to(cuda), torch.from_numpy(c).to(cuda))
    lab = lab.cpu().numpy()
    brute = ((x[:, None, :] - c[None, :, :]) ** 2).sum(2)
    # the oracle statement (sklearn's expansion) and the direct form agree except on sub-ulp ties
    ob = O.e_step(x, c)
    direct = brute.argmin(1)
    agree = ob == direct
    assert np.array_equal(lab[agree], ob[agree])
    assert lab[4] == 4 and lab[31] == 4
    # where the two fp64 formulations disagree the point is a true numerical tie: accept either
    assert np.all((lab[~agree] == ob[~agree]) | (lab[~agree] == direct[~agree]))


def test_assign_full_size_properties(cuda):
    """BASELINE config 2 at full size: 10 M rotations against K=1000.  Size-independent checks:
    a 200k-row sample against brute-force fp64, residual == x - c[label], idempotence, and the
    nearest key of every dictionary entry being itself."""
    from bdpose import ops
    N, K = 10_000_000, 1000
    g = torch.Generator(device=cuda).manual_seed(0)
    q = torch.randn(N, 4, device=cuda, generator=g, dtype=torch.float32)
    q = q / q.norm(dim=1, keepdim=True)
    w = q[:, :1].abs().clamp(max=1)
    x = (q[:, 1:] / q[:, 1:].norm(dim=1, keepdim=True) * (2 * torch.acos(w))).contiguous()
    c = x[:K].double().contiguous()
    lab, res, _ = ops.assign_nearest(x, c)
    assert int(lab.min()) >= 0 and int(lab.max()) < K
    assert torch.equal(lab[:K], torch.arange(K, device=cuda))
    assert torch.equal(res, (x.double() - c[lab]).float())
    lab2, _, _ = ops.assign_nearest(x, c, want_residual=False)
    assert torch.equal(lab, lab2)
    idx = torch.randperm(N, device=cuda, generator=g)[:200_000]
    xs = x[idx].double()
    best = torch.empty(idx.numel(), dtype=torch.int64, device=cuda)
    for s in range(0, idx.numel(), 20000):
        d = ((xs[s:s + 20000, None, :] - c[None, :, :]) ** 2).sum(2)
        best[s:s + 20000] = d.argmin(1)
    assert torch.equal(lab[idx], best)


# ---------------------------------------------------------------------------------------------------
# (c) k-means
# ---------------------------------------------------------------------------------------------------
def test_kmeans_golden(cuda, golden):
    from bdpose import kmeans
    g = golden("kmeans_fit")
    for sfx in ("", "_e"):
        r = kmeans.kmeans_lloyd(torch.from_numpy(g["X"]).to(cuda), torch.from_numpy(g["init" + sfx]).to(cuda))
        assert np.array_equal(r["labels"].cpu().numpy(), g["labels" + sfx]), sfx
        close(r["centers"], g["centers" + sfx], rtol=0, atol=1e-12)
        assert r["n_iter"] == int(g["n_iter" + sfx])
        close(r["inertia"], float(g["inertia" + sfx]), rtol=1e-12)


def test_kmeans_vs_sklearn_and_pickle(cuda, tmp_path):
    """Same explicit init -> same labels as scikit-learn (the reference's estimator); the fitted
    object pickles and predicts like the reference's dictionary file."""
    sk = pytest.importorskip("sklearn.cluster")
    from bdpose.kmeans import KMeans
    rng = np.random.default_rng(21)
    X = rand_rot(rng, 50_000)[0]
    K = 200
    init = X[:K].copy()
    ref = sk.KMeans(n_clusters=K, init=init, n_init=1, max_iter=25, tol=1e-4, algorithm="lloyd").fit(X)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ours = KMeans(n_clusters=K, init=init, max_iter=25, tol=1e-4, n_jobs=10).fit(X)
    assert ours.n_iter_ == ref.n_iter_
    assert np.array_equal(ours.labels_, ref.labels_)
    close(ours.cluster_centers_, ref.cluster_centers_, rtol=0, atol=1e-12)
    close(ours.inertia_, ref.inertia_, rtol=1e-11)
    f = tmp_path / "kmeans_dictionary_axis_angle_200.pkl"
    with open(f, "wb") as fh:
        pickle.dump(ours, fh)
    back = pickle.load(open(f, "rb"))
    assert back.n_clusters == K and back.cluster_centers_.dtype == np.float64
    y = rand_rot(rng, 12)[0].astype(np.float32)
    assert np.array_equal(back.predict(y), ref.predict(y.astype(np.float64)))


def test_kmeans_sharding_invariance(cuda):
    """The fixed-point accumulators make the centres independent of how rows are split: two half
    shards accumulated into one buffer give bit-identical centres to one full launch."""
    from bdpose import kmeans
    rng = np.random.default_rng(8)
    X = torch.from_numpy(rand_rot(rng, 200_000)[0]).to(cuda)
    K, d = 200, 3
    c = X[:K].clone().contiguous()
    hb = kmeans._fix_hi_bits(float(X.abs().max()))
    outs = []
    for parts in ([X], [X[:70_001].contiguous(), X[70_001:].contiguous()]):
        acc = None
        for part in parts:
            st = kmeans.LloydState(part.shape[0], K, d, cuda)
            kmeans.lloyd_step(part, c, st, hb, update=True)
            acc = st.acc_stats.clone() if acc is None else acc + st.acc_stats
        st.acc_stats.copy_(acc)
        new = torch.empty_like(c)
        kmeans.finalize(st, c, new, hb)
        outs.append(new.clone())
        assert int(acc[-2]) == X.shape[0]          # every label changed from -1
    assert torch.equal(outs[0], outs[1])
    # and the exact sums agree with a float64 reference to rounding
    lab = torch.from_numpy(O.e_step(X.cpu().numpy(), c.cpu().numpy())).long().to(cuda)
    ref = torch.zeros_like(c).index_add_(0, lab, X) / torch.bincount(lab, minlength=K).double()[:, None]
    close(outs[0], ref, rtol=0, atol=1e-12)


# ---------------------------------------------------------------------------------------------------
# (d) evaluation
# ---------------------------------------------------------------------------------------------------
def test_eval_golden(cuda, golden, capsys):
    import axisAngle
    import quaternion
    g = golden("eval_metrics")
    acc, med, err = axisAngle.get_error(g["gt"], g["hat"])
    assert "Error stats- Median:" in capsys.readouterr().out
    # fp64 inputs: same formula as the reference; differences are rounding noise amplified by acos
    close(err, g["err"], rtol=1e-9, atol=1e-6)
    assert acc == float(g["acc"])
    close(med, float(g["med"]), rtol=1e-9)
    close(axisAngle.get_error2(g["gt"], g["hat"], g["labels"], 12), float(g["e2"]), rtol=1e-9)
    accq, medq, errq = quaternion.get_error(g["gtq"], g["hatq"])
    close(errq, g["errq"], rtol=1e-9, atol=1e-6)
    assert accq == float(g["accq"])
    close(medq, float(g["medq"]), rtol=1e-9)
    close(quaternion.get_error2(g["gtq"], g["hatq"], g["labels"], 12), float(g["e2q"]), rtol=1e-9)
    assert np.isnan(axisAngle.get_error2(g["gt"], g["hat"], g["lab_missing"], 12))
    # float32 inputs are widened exactly
    _, _, e32 = axisAngle.get_error(g["gt"].astype(np.float32), g["hat"].astype(np.float32))
    close(e32, O.errors_aa(g["gt"].astype(np.float32).astype(np.float64),
                           g["hat"].astype(np.float32).astype(np.float64)), rtol=1e-9, atol=1e-6)


@pytest.mark.parametrize("N,C", [(1, 1), (2, 1), (1001, 1), (1000, 12), (100_003, 100), (4096, 7)])
def test_error_stats_vs_numpy(cuda, N, C):
    from bdpose import ops
    rng = np.random.default_rng(N + C)
    err = np.abs(rng.standard_normal(N)) * 40
    err[rng.integers(0, N, max(1, N // 10))] = err[0]         # repeated values
    labels = rng.integers(0, C, N)
    med, cnt, b30, mx = ops.error_stats(torch.from_numpy(err).to(cuda),
                                        torch.from_numpy(labels).to(cuda), C)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = np.array([np.median(err[labels == c]) for c in range(C)])
    np.testing.assert_array_equal(med.cpu().numpy(), ref)       # exact order statistics
    assert np.array_equal(cnt.cpu().numpy(), np.bincount(labels, minlength=C))
    assert int(b30) == int((err < 30).sum()) and float(mx) == err.max()


def test_eval_full_size_properties(cuda):
    """BASELINE config 5: 1 M predictions.  err(y, y) == 0, symmetry, range, and the device median /
    Acc@30 equal numpy's on the device-computed errors."""
    from bdpose import ops
    rng = np.random.default_rng(1)
    N = 1_000_000
    a = torch.from_numpy(rand_rot(rng, N)[0]).to(cuda)
    b = torch.from_numpy(rand_rot(rng, N)[0]).to(cuda)
    e = ops.geodesic_error_deg(a, b)
    assert float(ops.geodesic_error_deg(a, a).max()) < 1e-5
    close(ops.geodesic_error_deg(b, a), e, rtol=1e-9, atol=1e-6)
    assert float(e.min()) >= 0 and float(e.max()) <= 180.0 + 1e-9
    labels = torch.from_numpy(rng.integers(0, 12, N)).to(cuda)
    med, cnt, b30, mx = ops.error_stats(e, labels, 12)
    eh, lh = e.cpu().numpy(), labels.cpu().numpy()
    np.testing.assert_array_equal(med.cpu().numpy(), [np.median(eh[lh == c]) for c in range(12)])
    assert int(b30) == int((eh < 30).sum())
    sub = slice(0, 2000)
    close(e[sub], O.errors_aa(a[sub].cpu().numpy(), b[sub].cpu().numpy()), rtol=1e-9, atol=1e-6)


def test_kmeans_plusplus_fit_runs_and_improves(cuda):
    """The default path of learnKmeansDictionary.py:41-42 (k-means++ seeding, several restarts): the
    seeding stream is torch's, so there is no bit parity with sklearn here — check the fit contract
    (shapes, dtype, labels consistent with the centres, inertia no worse than a plain first-K init)."""
    from bdpose.kmeans import KMeans
    from bdpose import ops
    rng = np.random.default_rng(2)
    X = rand_rot(rng, 60_000)[0]
    K = 64
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        km = KMeans(n_clusters=K, n_init=2, max_iter=30, random_state=3, n_jobs=10).fit(X)
        base = KMeans(n_clusters=K, init=X[:K].copy(), max_iter=30).fit(X)
    assert km.cluster_centers_.shape == (K, 3) and km.cluster_centers_.dtype == np.float64
    assert km.labels_.shape == (X.shape[0],) and km.labels_.dtype == np.int32
    lab, _, sq = ops.assign_nearest(torch.from_numpy(X).to(cuda), torch.from_numpy(km.cluster_centers_).to(cuda),
                                    want_residual=False, want_sqdist=True, label_dtype=torch.int32)
    assert np.array_equal(lab.cpu().numpy(), km.labels_)
    close(float(sq.sum()), km.inertia_, rtol=1e-9)
    assert km.inertia_ <= base.inertia_ * 1.05
    assert np.array_equal(km.predict(X[:1000].astype(np.float32)), km.labels_[:1000])


def test_soft_bin_losses_golden(cuda, golden, tmp_path):
    """The soft-bin loss family (binDeltaLosses.py:109-208, 300-320) against the
    reference's own outputs: value and gradients w.r.t. score and residual.  These losses call
    `my_loss(ydata, centers[k] + residual)` with the prediction in the SECOND slot, so the pose
    loss has to propagate the gradient to its second argument."""
    import axisAngle
    import quaternion
    import binDeltaLosses as BL
    g = golden("losses_prob")
    kfile = tmp_path / "k.pkl"

    from bdpose.kmeans import KMeans
    km = KMeans(n_clusters=int(g["centers"].shape[0]))
    km.cluster_centers_ = np.asarray(g["centers"])
    with open(kfile, "wb") as f:
        pickle.dump(km, f)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(cuda)
    cases = {
        "prob": (lambda: BL.ProbabilisticLoss(0.7, str(kfile), axisAngle.geodesic_loss(reduce=False)), "res", "bins", "ydata"),
        "prob_multires": (lambda: BL.ProbabilisticMultiresLoss(0.7, str(kfile), axisAngle.geodesic_loss(reduce=False)), "res_k", "bins", "ydata"),
        "relaxed_q": (lambda: BL.RelaXedProbabilisticLossQ(0.7, str(kfile), quaternion.geodesic_loss(reduce=False)), "res_q", "prob", "ydata_q"),
        "m3_geo": (lambda: BL.loss_m3(0.7, str(kfile), axisAngle.geodesic_loss(reduce=False)), "res", "prob", "ydata"),
    }
    for fused in (True, False):
        BL.FUSE_EXPECTED_POSE = fused          # one fused launch / the reference's loop over bins
        for name, (make, rk, tk, yk) in cases.items():
            crit = make()
            s_ = t(g["score"]).requires_grad_(True)
            r_ = t(g[rk]).requires_grad_(True)
            loss = crit([s_, r_], [t(g[tk]), t(g[yk])])
            loss.backward()
            tag = name + (" fused" if fused else " loop")
            close(loss, g[name + "/loss"], rtol=1e-5, msg=tag + " loss")
            ref_s, ref_r = g[name + "/g_score"], g[name + "/g_res"]
            close(s_.grad, ref_s, rtol=0, atol=1e-5 * float(np.abs(ref_s).max()), msg=tag + " g_score")
            close(r_.grad, ref_r, rtol=0, atol=1e-5 * float(np.abs(ref_r).max()), msg=tag + " g_res")
    BL.FUSE_EXPECTED_POSE = True


def test_euler_to_pose_vs_oracle(cuda, golden):
    """Batched pose targets: Euler angles -> R -> axis-angle / quaternion against the
    oracle's restatement of helperFunctions.rotation_matrix + axisAngle.get_y / quaternion.get_y,
    on the golden Euler angles (whose matrices were produced by the reference) and on the special
    cases (identity -> zero vector, theta = pi -> zero axis-angle by the reference's convention)."""
    from bdpose import ops
    g = golden("rotation_helpers")
    eul = g["euler"]
    for i in range(len(eul)):                      # the oracle reproduces the reference's matrices
        assert np.allclose(O.rotation_matrix(*eul[i]), g["R_euler"][i], rtol=0, atol=1e-15)
    rng = np.random.default_rng(4)
    extra = np.concatenate([rng.uniform(-180, 360, (500, 3)),
                            [[0, 0, 0], [180, 0, 0], [0, 180, 0], [90, 0, -90], [30, 1e-7, -30]]])
    e = np.concatenate([eul, extra])
    aa, q = ops.euler_to_pose(torch.from_numpy(e).to(cuda), want_aa=True, want_quat=True)
    ref_aa = np.stack([O.get_y(O.rotation_matrix(*r)) for r in e])
    ref_q = np.stack([O.quat_get_y(O.rotation_matrix(*r)) for r in e])
    # near theta = pi the log map divides by a vanishing ||vee(R - R^T)||: compare where it is
    # well conditioned to 1e-9, everywhere to 1e-6
    n = np.linalg.norm(ref_aa, axis=1)
    ok = (n < 3.1) & ((n > 1e-3) | (n == 0))
    assert np.allclose(aa.cpu().numpy()[ok], ref_aa[ok], rtol=0, atol=1e-9)
    assert np.allclose(q.cpu().numpy()[ok], ref_q[ok], rtol=0, atol=1e-9)
    assert np.allclose(aa.cpu().numpy(), ref_aa, rtol=0, atol=1e-6)
    assert np.array_equal(aa.cpu().numpy()[len(eul) + 500], np.zeros(3))


def test_learn_dictionary_cli_synthetic(cuda, tmp_path):
    """bdpose.learn_dictionary (the GPU learnKmeansDictionary.py): fits, pickles an estimator that the
    loss / generator mirrors can load (n_clusters, cluster_centers_, predict)."""
    from bdpose import learn_dictionary
    out = tmp_path / "kmeans_dictionary_axis_angle_32.pkl"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        learn_dictionary.main(["32", "--synthetic", "40000", "--out", str(out), "--n_init", "2"])
    km = pickle.load(open(out, "rb"))
    assert km.n_clusters == 32 and km.cluster_centers_.shape == (32, 3)
    assert np.all(np.linalg.norm(km.cluster_centers_, axis=1) <= np.pi + 1e-9)
    assert km.predict(km.cluster_centers_.astype(np.float32)).tolist() == list(range(32))


@pytest.mark.parametrize("N,K,d,dt", [(5000, 1000, 3, np.float32), (777, 16, 4, np.float64),
                                      (300, 3000, 4, np.float32), (1, 1, 3, np.float32)])
def test_soft_assignment_vs_oracle(cuda, N, K, d, dt):
    """Row c4: p = exp(-g d^2) / sum, res = y - p @ C (binDeltaGenerators.py:104-108) against the
    numpy restatement; K = 3000, d = 4 takes the path that reads the dictionary from global memory."""
    from bdpose import ops
    rng = np.random.RandomState(N + K)
    c = rng.randn(K, d) * 0.7
    y = (c[rng.randint(0, K, N)] + 0.2 * rng.randn(N, d)).astype(dt)
    p_ref, r_ref = O.soft_assign(y, c, gamma=10.0)
    p, r = ops.assign_soft(torch.from_numpy(y).to(cuda), torch.from_numpy(c).to(cuda), 10.0)
    assert p.dtype == torch.float32 and r.dtype == torch.float32
    close(p, np.asarray(p_ref, dtype=np.float32), rtol=1e-5, atol=1e-30)
    close(r, np.asarray(r_ref, dtype=np.float32), rtol=1e-5, atol=1e-6)
    close(p.sum(1), np.ones(N), rtol=1e-5)
    # residual only / probabilities only
    p2, r2 = ops.assign_soft(torch.from_numpy(y).to(cuda), torch.from_numpy(c).to(cuda), 10.0, want_p=False)
    assert p2 is None and torch.equal(r2, r)
