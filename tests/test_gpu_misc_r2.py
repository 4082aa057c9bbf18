"""Round-2 small ops against golden vectors of the reference (tests/golden/make_golden.py misc):
loss_m2 (wrong gather axis), helperFunctions.get_gamma / mySGD / get_accuracy, the
test-time pose compositions of the scripts' testing() loops, the in-place
upstream-scalar path of the fused loss, and the MATLAB detection-metric helpers (hand-computed)."""
import numpy as np
import pytest
import torch

import bdpose_oracle as O

pytestmark = pytest.mark.gpu


def test_loss_m2_golden(cuda, golden):
    import binDeltaLosses as BL
    g = golden("misc_r2")
    s = torch.from_numpy(g["m2_score"]).to(cuda).requires_grad_(True)
    r = torch.from_numpy(g["m2_res"]).to(cuda).requires_grad_(True)
    crit = BL.loss_m2(float(g["m2_alpha"]), s.shape[1])
    loss = crit([s, r], [torch.from_numpy(g["m2_bins"]).to(cuda), torch.from_numpy(g["m2_res_true"]).to(cuda)])
    loss.backward()
    assert loss.dim() == 0
    np.testing.assert_allclose(loss.item(), g["m2_loss"], rtol=1e-5)
    np.testing.assert_allclose(s.grad.cpu().numpy(), g["m2_g_score"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(r.grad.cpu().numpy(), g["m2_g_res"], rtol=1e-5, atol=1e-8)


def test_get_gamma_and_accuracy_golden(cuda, golden):
    import helperFunctions as H
    g = golden("misc_r2")
    assert H.get_gamma(g["gamma_centers"]) == pytest.approx(float(g["gamma"]), rel=1e-12)
    assert H.get_accuracy(g["acc_true"], g["acc_pred"], 5) == float(g["acc"])
    assert H.parse_name("n0123_m07_a12.5_e-3.25_t0.5_d2.0") == ("n0123", "m07", 12.5, -3.25, 0.5, 2.0)
    np.testing.assert_allclose(H.rotation_matrix(30.0, -10.0, 5.0), O.rotation_matrix(30.0, -10.0, 5.0), atol=1e-15)
    assert H.eps == 1e-6 and len(H.classes) == 12


def test_compose_prediction_golden(cuda, golden):
    """dict[argmax] (+) residual in the three forms of the scripts' testing() loops: outputs equal
    the numpy lines of the scripts (5e-6 absolute for the Riemannian form, whose script evaluates
    get_R on float32 residuals), bins bit-exact incl. the planted argmax tie."""
    from bdpose import ops
    g = golden("misc_r2")
    sc = torch.from_numpy(g["t_score"]).to(cuda)
    ref_bin = np.argmax(g["t_score"], axis=1)
    y, b = ops.compose_prediction(sc, torch.from_numpy(g["t_res"]).to(cuda), torch.from_numpy(g["t_dict"]).to(cuda))
    assert y.dtype == torch.float64 and np.array_equal(b.cpu().numpy(), ref_bin)
    np.testing.assert_array_equal(y.cpu().numpy(), g["t_add"])
    y, _ = ops.compose_prediction(sc, torch.from_numpy(g["t_res4"]).to(cuda), torch.from_numpy(g["t_qdict"]).to(cuda),
                                  mode="normalize")
    np.testing.assert_allclose(y.cpu().numpy(), g["t_quat"], rtol=0, atol=1e-15)
    y, _ = ops.compose_prediction(sc, torch.from_numpy(g["t_res"]).to(cuda),
                                  torch.from_numpy(g["t_rotdict"]).to(cuda), mode="riemannian")
    # (the script line runs get_R on float32 residuals: its own rounding is ~1e-7 relative, amplified
    # near theta = pi by the log map — 5e-6 absolute against it, 1e-12 against the fp64 restatement)
    np.testing.assert_allclose(y.cpu().numpy(), g["t_riem"], rtol=0, atol=5e-6)
    # the fp64 oracle restatement (same arithmetic precision as the kernel) to 1e-12
    np.testing.assert_allclose(y.cpu().numpy(), O.compose_riemannian(g["t_score"], g["t_res"].astype(np.float64),
                                                                     g["t_rotdict"]), rtol=0, atol=1e-12)
    # row pitch (logits that are a column window) and a large ragged batch
    big = torch.randn(100_003, 200 + 8, device=cuda)
    res = torch.randn(100_003, 3, device=cuda) * 0.1
    dic = torch.randn(200, 3, device=cuda, dtype=torch.float64)
    y, b = ops.compose_prediction(big[:, :200], res, dic)
    rb = torch.argmax(big[:, :200], dim=1)
    assert torch.equal(b, rb)
    assert torch.equal(y, dic[rb] + res.double())


@pytest.mark.parametrize("name,kw", [("plain", {}), ("mom", dict(momentum=0.9, weight_decay=1e-2)),
                                     ("nest", dict(momentum=0.8, nesterov=True, dampening=0.0))])
def test_mysgd_golden(cuda, golden, name, kw):
    """helperFunctions.mySGD: 7 steps of the cyclical learning rate, every parameter in one launch."""
    import helperFunctions as H
    g = golden("misc_r2")
    q1 = torch.from_numpy(g["sgd_p1"].copy()).to(cuda).requires_grad_(True)
    q2 = torch.from_numpy(g["sgd_p2"].copy()).to(cuda).requires_grad_(True)
    t1, t2 = torch.from_numpy(g["sgd_t1"]).to(cuda), torch.from_numpy(g["sgd_t2"]).to(cuda)
    opt = H.mySGD([q1, q2], c=4, alpha1=1e-1, alpha2=1e-3, **kw)
    for it in range(7):
        opt.zero_grad()
        (((q1 - t1) ** 2).sum() + ((q2 - t2) ** 4).sum()).backward()
        opt.step()
        cur = np.concatenate([q1.detach().cpu().numpy().ravel(), q2.detach().cpu().numpy().ravel()])
        np.testing.assert_allclose(cur, g["sgd_traj_" + name][it], rtol=2e-5, atol=1e-6,
                                   err_msg="%s step %d" % (name, it))
    assert opt.state[q1]["step"] == 7


def test_bd_loss_upstream_scalars_in_place(cuda):
    """loss = Lc + w * Lr (learnGeodesicBDModel.py:180,185; learnObjectnetBDModel.py:140): the
    gradients equal the oracle's for w != 1 and for an upstream factor on Lc, and the logits gradient
    is the buffer the forward launch wrote (scaled in place: no second [B, K] tensor)."""
    from bdpose import ops, _lib as L
    torch.manual_seed(3)
    B, K = 257, 40
    score, delta = torch.randn(B, K), torch.randn(B, 3) * 0.2
    bins = torch.randint(0, K, (B,))
    target, keys = torch.randn(B, 3), torch.randn(K, 3)
    for wc, wr in ((1.0, 0.37), (0.1, 10.0)):
        s1 = score.clone().requires_grad_(True); d1 = delta.clone().requires_grad_(True)
        l1, l2 = O.bin_delta_terms(s1, d1, bins, target, keys, "aa")
        (wc * l1 + wr * l2).backward()
        s2 = score.to(cuda).requires_grad_(True); d2 = delta.to(cuda).requires_grad_(True)
        lc, lr, _ = ops.bd_loss(s2, bins.to(cuda), d2, target.to(cuda), keys.to(cuda), L.POSE_GEODESIC_AA, True)
        (wc * lc + wr * lr).backward()
        assert torch.allclose(s2.grad.cpu(), s1.grad, rtol=1e-4, atol=1e-8)
        assert torch.allclose(d2.grad.cpu(), d1.grad, rtol=1e-4, atol=1e-8)


def test_detection_metric_helpers():
    """box_overlap.m / VOCap.m ports on hand-computed cases."""
    from bdpose import metrics
    o = metrics.box_overlap(np.array([[0, 0, 9, 9], [5, 5, 14, 14], [20, 20, 30, 30]]), np.array([0, 0, 9, 9]))
    np.testing.assert_allclose(o, [1.0, 25.0 / 175.0, 0.0])
    assert metrics.VOCap([0.5, 1.0], [1.0, 0.5]) == pytest.approx(0.75)
    assert metrics.VOCap([0.2, 0.2, 0.6], [1.0, 0.5, 0.6]) == pytest.approx(0.2 * 1.0 + 0.4 * 0.6)
