"""The oracle (oracle/bdpose_oracle.py) against the golden vectors produced by the reference's own
modules (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

import bdpose_oracle as O


def test_rotation_helpers(golden):
    g = golden("rotation_helpers")
    for v, R, y, q in zip(g["aa"], g["R"], g["y"], g["q"]):
        np.testing.assert_allclose(O.get_R(v), R, rtol=0, atol=1e-15)
        np.testing.assert_allclose(O.get_y(R), y, rtol=0, atol=1e-15)
        np.testing.assert_allclose(O.quat_get_y(R), q, rtol=0, atol=1e-15)
    np.testing.assert_allclose(O.convert_dictionary(g["aa"]), g["qdict"], rtol=0, atol=1e-15)
    for e, R in zip(g["euler"], g["R_euler"]):
        np.testing.assert_allclose(O.rotation_matrix(*e), R, rtol=0, atol=1e-15)


def test_eval_metrics(golden):
    g = golden("eval_metrics")
    acc, med, err = O.get_error(g["gt"], g["hat"])
    np.testing.assert_array_equal(err, g["err"])
    assert acc == float(g["acc"]) and med == float(g["med"])
    assert O.get_error2(g["gt"], g["hat"], g["labels"], 12) == float(g["e2"])
    accq, medq, errq = O.get_error(g["gtq"], g["hatq"], quaternion=True)
    np.testing.assert_array_equal(errq, g["errq"])
    assert accq == float(g["accq"]) and medq == float(g["medq"])
    assert O.get_error2(g["gtq"], g["hatq"], g["labels"], 12, quaternion=True) == float(g["e2q"])
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.isnan(O.get_error2(g["gt"], g["hat"], g["lab_missing"], 12))
    assert np.isnan(float(g["e2_nan"]))


def _grad(fn, *leaves):
    leaves = [torch.from_numpy(np.array(a)).requires_grad_(True) for a in leaves]
    out = fn(*leaves)
    out.backward()
    return out.detach().numpy(), [t.grad.numpy() for t in leaves]


def test_losses(golden):
    g = golden("losses")
    T = lambda k: torch.from_numpy(g[k])
    alpha = float(g["alpha"])
    # stand-alone pose losses
    v, (gp,) = _grad(lambda p: O.geodesic_loss_aa(p, T("ytrue_aa")), g["p_aa"])
    np.testing.assert_allclose(v, g["geo_aa_loss"], rtol=1e-6)
    np.testing.assert_allclose(gp, g["geo_aa_g0"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(O.geodesic_loss_aa(T("p_aa"), T("ytrue_aa"), reduce=False).numpy(),
                               g["geo_aa_rows"], rtol=1e-6)
    v, (gp,) = _grad(lambda p: O.geodesic_loss_quat(p, T("ytrue_q")), g["p_q"])
    np.testing.assert_allclose(v, g["geo_q_loss"], rtol=1e-6)
    np.testing.assert_allclose(gp, g["geo_q_g0"], rtol=1e-5, atol=1e-8)
    # composites
    centers = torch.from_numpy(g["centers"]).float()
    qcenters = torch.from_numpy(O.convert_dictionary(g["centers"])).float()
    cases = [
        ("simple", "res3", dict(target=T("res_true"), centers=None, pose="mse")),
        ("geod_mse", "res3", dict(target=T("ytrue_aa"), centers=centers, pose="mse")),
        ("geod_aa", "res3", dict(target=T("ytrue_aa"), centers=centers, pose="aa")),
        ("geod_q", "res4", dict(target=T("ytrue_q"), centers=qcenters, pose="quat")),
    ]
    for name, rk, kw in cases:
        def f(s, r):
            l1, l2 = O.bin_delta_terms(s, r, T("bin_true"), **kw)
            return l1 + alpha * l2
        v, (gs, gr) = _grad(f, g["score"], g[rk])
        np.testing.assert_allclose(v, g[name + "_loss"], rtol=1e-6, err_msg=name)
        np.testing.assert_allclose(gs, g[name + "_g0"], rtol=1e-5, atol=1e-8, err_msg=name)
        np.testing.assert_allclose(gr, g[name + "_g1"], rtol=1e-5, atol=1e-8, err_msg=name)

    def f(s, r):
        l1, l2 = O.riemannian_terms(s, r, T("bin_true"), T("R_true"), T("key_rot").float())
        return l1 + alpha * l2
    v, (gs, gr) = _grad(f, g["score"], g["res3"])
    np.testing.assert_allclose(v, g["riem_loss"], rtol=1e-6)
    np.testing.assert_allclose(gs, g["riem_g0"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(gr, g["riem_g1"], rtol=1e-5, atol=1e-8)


def test_label_generation(golden):
    g = golden("label_generation")
    c = g["centers"]
    ya = g["ydata_aa"].reshape(-1, 3)
    b, r = O.predict_residual(ya, c)
    np.testing.assert_array_equal(b, g["gbd_bin"].reshape(-1))
    np.testing.assert_array_equal(r, g["gbd_res"].reshape(-1, 3))
    qc = O.convert_dictionary(c)
    yq = g["ydata_q"].reshape(-1, 4)
    b, r = O.predict_residual(yq, qc)
    np.testing.assert_array_equal(b, g["gbdq_bin"].reshape(-1))
    np.testing.assert_array_equal(r, g["gbdq_res"].reshape(-1, 4))
    p, r = O.soft_assign(yq, qc)
    np.testing.assert_allclose(p, g["xpbdq_bin"].reshape(-1, c.shape[0]), rtol=1e-6, atol=1e-30)
    np.testing.assert_allclose(r, g["xpbdq_res"].reshape(-1, 4), rtol=1e-6, atol=1e-7)
    b, r, rot = O.riemannian_targets(ya, c)
    np.testing.assert_array_equal(b, g["rbd_bin"].reshape(-1))
    np.testing.assert_array_equal(r, g["rbd_res"].reshape(-1, 3))
    np.testing.assert_array_equal(rot, g["rbd_rot"].reshape(-1, 3, 3))
    b, r = O.quatdot_assign(g["qq"], g["qkeys"])
    np.testing.assert_array_equal(b, g["qbin"])
    np.testing.assert_array_equal(r, g["qres"])


def test_kmeans_fit(golden):
    g = golden("kmeans_fit")
    for sfx in ("", "_e"):
        r = O.kmeans_lloyd(g["X"], g["init" + sfx])
        np.testing.assert_array_equal(r["labels"], g["labels" + sfx])
        np.testing.assert_allclose(r["centers"], g["centers" + sfx], rtol=0, atol=1e-12)
        assert r["n_iter"] == int(g["n_iter" + sfx])
        np.testing.assert_allclose(r["inertia"], float(g["inertia" + sfx]), rtol=1e-12)


def _load_heads(g):
    C, K, N0, N1, N2, nd, B = [int(v) for v in g["dims"]]
    m = O.OneBinDeltaHeads(C, K, N0, N1, N2, nd)
    sd = {k[len("sd0/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd0/")}
    m.load_state_dict(sd)
    return m


def test_heads(golden):
    g = golden("heads")
    m = _load_heads(g)
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    label = torch.from_numpy(g["label"])
    m.train()
    y1, y2 = m(x, label)
    np.testing.assert_allclose(y1.detach().numpy(), g["train_y1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(y2.detach().numpy(), g["train_y2"], rtol=1e-5, atol=1e-6)
    ((y1 * torch.from_numpy(g["w1"])).sum() + (y2 * torch.from_numpy(g["w2"])).sum()).backward()
    np.testing.assert_allclose(x.grad.numpy(), g["train_gx"], rtol=1e-4, atol=1e-6)
    for k, p in m.named_parameters():
        np.testing.assert_allclose(p.grad.numpy(), g["train_grad/" + k], rtol=1e-4, atol=1e-6,
                                   err_msg=k)
    for k, v in m.state_dict().items():
        if "running" in k:
            np.testing.assert_allclose(v.numpy(), g["train_sd/" + k], rtol=1e-6, err_msg=k)
    m.eval()
    with torch.no_grad():
        e1, e2 = m(x, label)
    np.testing.assert_allclose(e1.numpy(), g["eval_y1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(e2.numpy(), g["eval_y2"], rtol=1e-5, atol=1e-6)


def test_per_bin_delta_heads_oracle_vs_reference(golden):
    """oracle.OneDeltaPerBinHeads reproduces the reference's OneDeltaPerBinModel /
    ProbabilisticOneDeltaPerBinModel (binDeltaModels.py:124-178) from heads_perbin.npz."""
    import torch
    import bdpose_oracle as O
    g = golden("heads_perbin")
    Cc, Kc, N0, N1, N2, N3, nd, Bh = [int(v) for v in g["dims"]]
    m = O.OneDeltaPerBinHeads(Cc, Kc, N0, N1, N2, N3, nd)
    m.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd0/")})
    x, lab = torch.from_numpy(g["x"]), torch.from_numpy(g["label"])
    m.train()
    xr = x.clone().requires_grad_(True)
    y1, y2 = m(xr, lab)
    (y1 * torch.from_numpy(g["w1"])).sum().add((y2 * torch.from_numpy(g["w2"])).sum()).backward()
    assert torch.allclose(y1, torch.from_numpy(g["train_y1"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(y2, torch.from_numpy(g["train_y2"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(xr.grad, torch.from_numpy(g["train_gx"]), rtol=1e-4, atol=1e-6)
    m.eval()
    with torch.no_grad():
        e1, e2 = m(x, lab)
        p1, p2 = m(x, lab, probabilistic=True)
    assert torch.allclose(e2, torch.from_numpy(g["eval_y2"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(p2, torch.from_numpy(g["prob_y2"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(p1, torch.from_numpy(g["prob_y1"]), rtol=1e-5, atol=1e-6)


def test_round2_misc(golden):
    """loss_m2, get_gamma, the testing() compositions, mySGD and get_accuracy of the reference
    (tests/golden/make_golden.py misc)."""
    g = golden("misc_r2")
    s = torch.from_numpy(g["m2_score"]).requires_grad_(True)
    r = torch.from_numpy(g["m2_res"]).requires_grad_(True)
    loss = O.loss_m2(s, r, torch.from_numpy(g["m2_bins"]), torch.from_numpy(g["m2_res_true"]), float(g["m2_alpha"]))
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["m2_loss"], rtol=1e-6)
    np.testing.assert_allclose(s.grad.numpy(), g["m2_g_score"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(r.grad.numpy(), g["m2_g_res"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(O.get_gamma(g["gamma_centers"]), g["gamma"], rtol=1e-12)
    np.testing.assert_array_equal(O.compose_add(g["t_score"], g["t_res"], g["t_dict"]), g["t_add"])
    np.testing.assert_array_equal(O.compose_normalize(g["t_score"], g["t_res4"], g["t_qdict"]), g["t_quat"])
    np.testing.assert_allclose(O.compose_riemannian(g["t_score"], g["t_res"], g["t_rotdict"]), g["t_riem"],
                               rtol=0, atol=5e-6)      # the script evaluates get_R on float32 residuals
    t1, t2 = g["sgd_t1"], g["sgd_t2"]

    def grads(p):
        return [2 * (p[0] - t1), 4 * (p[1] - t2) ** 3]
    for name, kw in (("plain", {}), ("mom", dict(momentum=0.9, weight_decay=1e-2)),
                     ("nest", dict(momentum=0.8, nesterov=True))):
        traj = O.my_sgd([g["sgd_p1"], g["sgd_p2"]], grads, 7, 4, 1e-1, 1e-3, **kw)
        np.testing.assert_allclose(traj, g["sgd_traj_" + name], rtol=2e-5, atol=1e-6)
    assert O.get_accuracy(g["acc_true"], g["acc_pred"], 5) == float(g["acc"])
