"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN MODULES (read-only checkout at
/root/reference) on seeded inputs.  Run once in the build container:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so the vectors are committed.  Shims applied (all kept
outside the reference tree): `.cuda()` is a no-op (no GPU here); torchvision's
pretrained ResNet download is replaced by an identity feature model; sklearn's `predict` input is
cast to float64 and `n_features_in_` follows the swapped-in quaternion dictionary (sklearn >= 1.0
rejects float32 X against float64 centres and checks the feature count; the reference targeted
sklearn ~0.19); `KMeans(n_jobs=...)` kwarg dropped.
"""
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch
from torch import nn

REF = os.environ.get("BDP_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    for name in ("tensorboardX", "progressbar"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)


def rand_rotations(rng, n):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 0] < 0] *= -1
    ang = 2 * np.arccos(np.clip(q[:, 0], -1, 1))
    ax = q[:, 1:] / np.maximum(np.linalg.norm(q[:, 1:], axis=1, keepdims=True), 1e-300)
    return ax * ang[:, None], q


def main():
    install_shims()
    import axisAngle as ref_aa
    import quaternion as ref_q
    import binDeltaLosses as ref_losses
    import binDeltaModels as ref_models
    import helperFunctions as ref_h
    from sklearn.cluster import KMeans

    rng = np.random.default_rng(0)
    torch.manual_seed(0)

    # ---- 1. rotation helpers ------------------------------------------------------------------
    aa, quat = rand_rotations(rng, 64)
    aa[0] = 0.0                                 # theta < eps -> identity (axisAngle.py:35)
    aa[1] = np.array([np.pi, 0, 0])             # theta = pi -> get_y returns 0 (axisAngle.py:24-27)
    aa[2] = np.array([1e-7, 0, 0])
    Rs = np.stack([ref_aa.get_R(v) for v in aa])
    ys = np.stack([ref_aa.get_y(R) for R in Rs])
    qs = np.stack([ref_q.get_y(R) for R in Rs])
    qdict = ref_q.convert_dictionary(aa)
    eul = rng.uniform(-180, 180, (16, 3))
    Reul = np.stack([ref_h.rotation_matrix(*e) for e in eul])
    np.savez(os.path.join(OUT, "rotation_helpers.npz"), aa=aa, R=Rs, y=ys, q=qs, qdict=qdict,
             euler=eul, R_euler=Reul)

    # ---- 2. evaluation metrics ----------------------------------------------------------------
    N = 600
    gt, gtq = rand_rotations(rng, N)
    hat, hatq = rand_rotations(rng, N)
    hat[:50] = gt[:50] + 0.05 * rng.standard_normal((50, 3))      # small errors too
    hat[50] = gt[50]                                              # exact zero error
    hatq[:50] = gtq[:50] + 0.02 * rng.standard_normal((50, 4))
    hatq[:50] /= np.linalg.norm(hatq[:50], axis=1, keepdims=True)
    hatq[50] = gtq[50]
    labels = rng.integers(0, 12, (N, 1))
    acc, med, err = ref_aa.get_error(gt, hat)
    e2 = ref_aa.get_error2(gt, hat, labels, 12)
    accq, medq, errq = ref_q.get_error(gtq, hatq)
    e2q = ref_q.get_error2(gtq, hatq, labels, 12)
    lab_missing = labels.copy()
    lab_missing[lab_missing == 7] = 3                              # empty class -> NaN median
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            e2_nan = ref_aa.get_error2(gt, hat, lab_missing, 12)
    np.savez(os.path.join(OUT, "eval_metrics.npz"), gt=gt, hat=hat, gtq=gtq, hatq=hatq, labels=labels,
             acc=acc, med=med, err=err, e2=e2, accq=accq, medq=medq, errq=errq, e2q=e2q,
             lab_missing=lab_missing, e2_nan=e2_nan)

    # ---- 3. losses ----------------------------------------------------------------------------
    B, K = 37, 24
    centers, _ = rand_rotations(rng, K)
    km = types.SimpleNamespace(cluster_centers_=centers.copy(), n_clusters=K)
    kfile = os.path.join(tempfile.mkdtemp(), "km.pkl")
    with open(kfile, "wb") as f:
        pickle.dump(_PickleDict(centers), f)
    score = torch.randn(B, K) * 2
    score[3, 5] = score[3, 9] = score[3].max() + 1.0               # argmax tie -> lowest index
    bin_true = torch.randint(0, K, (B,))
    ytrue_aa = torch.from_numpy(rand_rotations(rng, B)[0]).float()
    ytrue_q = torch.from_numpy(rand_rotations(rng, B)[1]).float()
    res3 = torch.randn(B, 3) * 0.3
    res3[0] = 0.0                                                  # zero residual row
    res4 = torch.randn(B, 4) * 0.3
    out = dict(score=score.numpy(), bin_true=bin_true.numpy(), ytrue_aa=ytrue_aa.numpy(),
               ytrue_q=ytrue_q.numpy(), res3=res3.numpy(), res4=res4.numpy(), centers=centers)

    def run(name, crit, ypred, ytrue):
        leaves = [t.clone().requires_grad_(True) for t in ypred]
        loss = crit(leaves, ytrue)
        loss.backward()
        out[name + "_loss"] = loss.detach().numpy()
        for i, t in enumerate(leaves):
            out["%s_g%d" % (name, i)] = t.grad.numpy() if t.grad is not None else np.zeros(0)

    # stand-alone pose losses (value + gradient), including a clamp-saturated row
    p_aa = (ytrue_aa + 0.2 * torch.randn(B, 3)).clone()
    p_aa[1] = ytrue_aa[1]                                          # |w| -> 1: clamp saturates
    out["p_aa"] = p_aa.numpy()
    run("geo_aa", lambda yp, yt: ref_aa.geodesic_loss()(yp[0], yt), [p_aa], ytrue_aa)
    out["geo_aa_rows"] = ref_aa.geodesic_loss(reduce=False)(p_aa, ytrue_aa).numpy()
    p_q = (ytrue_q + 0.2 * torch.randn(B, 4)).clone()
    p_q[1] = ytrue_q[1] * 1.7
    out["p_q"] = p_q.numpy()
    run("geo_q", lambda yp, yt: ref_q.geodesic_loss()(yp[0], yt), [p_q], ytrue_q)
    out["geo_q_rows"] = ref_q.geodesic_loss(reduce=False)(p_q, ytrue_q).numpy()

    alpha = 0.7
    res_true = torch.randn(B, 3) * 0.2
    out["res_true"] = res_true.numpy()
    out["alpha"] = alpha
    run("simple", ref_losses.SimpleLoss(alpha), [score, res3], [bin_true, res_true])
    run("geod_mse", ref_losses.GeodesicLoss(alpha, kfile), [score, res3], [bin_true, ytrue_aa])
    run("geod_aa", ref_losses.GeodesicLoss(alpha, kfile, ref_aa.geodesic_loss()), [score, res3],
        [bin_true, ytrue_aa])
    run("geod_q", ref_losses.GeodesicLossQ(alpha, kfile, ref_q.geodesic_loss()), [score, res4],
        [bin_true, ytrue_q])
    key_rot = np.stack([ref_aa.get_R(c) for c in centers])
    R_true = torch.from_numpy(np.stack([ref_aa.get_R(v) for v in ytrue_aa.numpy().astype(np.float64)])
                              ).float()
    out["key_rot"] = key_rot
    out["R_true"] = R_true.numpy()
    run("riem", ref_losses.RiemannianLoss(alpha, key_rot), [score, res3], [bin_true, R_true])
    np.savez(os.path.join(OUT, "losses.npz"), **out)

    # ---- 4. label generation (the reference's __getitem__ bodies on a synthetic ImagesAll) ------
    import dataGenerators as ref_dg
    import binDeltaGenerators as ref_gen
    from sklearn.cluster import KMeans as SKK
    Kd = 40
    train, _ = rand_rotations(rng, 4000)
    init = train[:Kd].copy()
    sk = SKK(n_clusters=Kd, init=init, n_init=1, max_iter=300, tol=1e-4, algorithm="lloyd").fit(train)
    orig_predict = SKK.predict

    def predict64(self, X):
        # GBDGeneratorQ swaps in a [K,4] quaternion dictionary (binDeltaGenerators.py:67); old sklearn
        # did not track n_features_in_
        X = np.asarray(X, dtype=np.float64)
        self.n_features_in_ = X.shape[1]
        return orig_predict(self, X)

    SKK.predict = predict64
    dfile = os.path.join(tempfile.mkdtemp(), "dict.pkl")
    with open(dfile, "wb") as f:
        pickle.dump(sk, f)
    n_items, C = 20, 12
    ydata_aa = rand_rotations(rng, n_items * C)[0].astype(np.float32).reshape(n_items, C, 3)
    ydata_q = rand_rotations(rng, n_items * C)[1].astype(np.float32).reshape(n_items, C, 4)

    def fake_init(self, db_path, db_type, ydata_type="axis_angle"):
        self.num_images = np.array([n_items] * C)
        self.ydata_type = ydata_type

    def fake_getitem(self, idx):
        y = ydata_aa[idx] if self.ydata_type == "axis_angle" else ydata_q[idx]
        return {"ydata": torch.from_numpy(y.copy()).float()}

    ref_dg.ImagesAll.__init__ = fake_init
    ref_dg.ImagesAll.__getitem__ = fake_getitem
    gen = dict(centers=sk.cluster_centers_, ydata_aa=ydata_aa, ydata_q=ydata_q)
    for cls, key in ((ref_gen.GBDGenerator, "gbd"), (ref_gen.GBDGeneratorQ, "gbdq"),
                     (ref_gen.XPBDGeneratorQ, "xpbdq"), (ref_gen.RBDGenerator, "rbd")):
        g = cls("unused", "real", dfile)
        items = [g[i] for i in range(n_items)]
        gen[key + "_bin"] = np.stack([s["ydata_bin"].numpy() for s in items])
        gen[key + "_res"] = np.stack([s["ydata_res"].numpy() for s in items])
        if "ydata_rot" in items[0]:
            gen[key + "_rot"] = np.stack([s["ydata_rot"].numpy() for s in items])
    # learnObjectnetModel.py:60-66, 108-109 (quaternion-dot assignment against the fixed 16 keys)
    s = 1 / np.sqrt(2)
    qkeys = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [s, s, 0, 0], [s, 0, s, 0],
                      [s, 0, 0, s], [0, s, s, 0], [0, s, 0, s], [0, 0, s, s], [s, -s, 0, 0],
                      [s, 0, -s, 0], [s, 0, 0, -s], [0, s, -s, 0], [0, s, 0, -s], [0, 0, s, -s]])
    qq = rand_rotations(rng, 300)[1]
    qbin = np.array([np.argmax(np.abs(np.dot(qkeys, t))) for t in qq])
    qres = np.stack([(t - qkeys[b, :]) for t, b in zip(qq, qbin)]).astype(np.float32)
    gen.update(qkeys=qkeys, qq=qq, qbin=qbin, qres=qres)
    np.savez(os.path.join(OUT, "label_generation.npz"), **gen)

    # ---- 5. k-means fit (learnKmeansDictionary.py:41-42 with an explicit init, n_init=1) --------
    X, _ = rand_rotations(rng, 6000)
    Kf = 25
    initf = X[:Kf].copy()
    skf = SKK(n_clusters=Kf, init=initf, n_init=1, max_iter=300, tol=1e-4, algorithm="lloyd").fit(X)
    # a run that must relocate an empty cluster: one init centre far from every sample
    init_e = initf.copy()
    init_e[3] = np.array([50.0, 50.0, 50.0])
    ske = SKK(n_clusters=Kf, init=init_e, n_init=1, max_iter=300, tol=1e-4, algorithm="lloyd").fit(X)
    np.savez(os.path.join(OUT, "kmeans_fit.npz"), X=X, init=initf, centers=skf.cluster_centers_,
             labels=skf.labels_, inertia=skf.inertia_, n_iter=skf.n_iter_, init_e=init_e,
             centers_e=ske.cluster_centers_, labels_e=ske.labels_, inertia_e=ske.inertia_,
             n_iter_e=ske.n_iter_)

    # ---- 6. heads (binDeltaModels.py:62-121) on small layer sizes --------------------------------
    import featureModels
    ident = lambda *a, **k: nn.Identity()
    ref_models.resnet_model = ident
    torch.manual_seed(1)
    Cc, Kc, N0, N1, N2, nd, Bh = 3, 16, 64, 40, 24, 3, 10
    model = ref_models.OneBinDeltaModel("resnet", Cc, Kc, N0, N1, N2, nd)
    # non-trivial BN affine parameters and running stats
    for m in model.modules():
        if isinstance(m, nn.BatchNorm1d):
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.uniform_(-0.3, 0.3)
            m.running_mean.uniform_(-0.2, 0.2)
            m.running_var.uniform_(0.5, 2.0)
    sd0 = {k: v.clone().numpy() for k, v in model.state_dict().items()}
    x = torch.randn(Bh, N0)
    label = torch.randint(0, Cc, (Bh, 1))
    head = dict(x=x.numpy(), label=label.numpy(), dims=np.array([Cc, Kc, N0, N1, N2, nd, Bh]))
    for k, v in sd0.items():
        head["sd0/" + k] = v
    model.train()
    xr = x.clone().requires_grad_(True)
    y1, y2 = model(xr, label)
    w1 = torch.randn_like(y1)
    w2 = torch.randn_like(y2)
    (y1 * w1).sum().add((y2 * w2).sum()).backward()
    head.update(train_y1=y1.detach().numpy(), train_y2=y2.detach().numpy(), w1=w1.numpy(),
                w2=w2.numpy(), train_gx=xr.grad.numpy())
    for k, p in model.named_parameters():
        head["train_grad/" + k] = p.grad.numpy()
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            head["train_sd/" + k] = v.clone().numpy()
    model.eval()
    with torch.no_grad():
        e1, e2_ = model(x, label)
    head.update(eval_y1=e1.numpy(), eval_y2=e2_.numpy())
    np.savez(os.path.join(OUT, "heads.npz"), **head)
    print("golden vectors written to", OUT)


def main_perbin():
    """heads_perbin.npz: OneDeltaPerBinModel / ProbabilisticOneDeltaPerBinModel of the reference
    (binDeltaModels.py:124-178) on small layer sizes; own seeds, so the other files are untouched.
    Run with `python tests/golden/make_golden.py perbin`."""
    install_shims()
    import binDeltaModels as ref_models
    ref_models.resnet_model = lambda *a, **k: nn.Identity()
    torch.manual_seed(11)
    Cc, Kc, N0, N1, N2, N3, nd, Bh = 3, 4, 64, 40, 24, 12, 3, 10
    model = ref_models.OneDeltaPerBinModel("resnet", Cc, Kc, N0, N1, N2, N3, nd)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm1d):
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.uniform_(-0.3, 0.3)
            m.running_mean.uniform_(-0.2, 0.2)
            m.running_var.uniform_(0.5, 2.0)
    out = dict(dims=np.array([Cc, Kc, N0, N1, N2, N3, nd, Bh]))
    for k, v in model.state_dict().items():
        out["sd0/" + k] = v.clone().numpy()
    x = torch.randn(Bh, N0)
    label = torch.randint(0, Cc, (Bh, 1))
    out.update(x=x.numpy(), label=label.numpy())
    model.train()
    xr = x.clone().requires_grad_(True)
    y1, y2 = model(xr, label)
    w1, w2 = torch.randn_like(y1), torch.randn_like(y2)
    (y1 * w1).sum().add((y2 * w2).sum()).backward()
    out.update(train_y1=y1.detach().numpy(), train_y2=y2.detach().numpy(), w1=w1.numpy(), w2=w2.numpy(),
               train_gx=xr.grad.numpy())
    for k, p in model.named_parameters():
        out["train_grad/" + k] = p.grad.numpy()
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["train_sd/" + k] = v.clone().numpy()
    # the probabilistic model shares the layer structure: same weights, all K deltas of the class
    prob = ref_models.ProbabilisticOneDeltaPerBinModel("resnet", Cc, Kc, N0, N1, N2, N3, nd)
    prob.load_state_dict(model.state_dict())
    prob.eval()
    model.eval()
    with torch.no_grad():
        e1, e2_ = model(x, label)
        p1, p2 = prob(x, label)
    out.update(eval_y1=e1.numpy(), eval_y2=e2_.numpy(), prob_y1=p1.numpy(), prob_y2=p2.numpy())
    np.savez(os.path.join(OUT, "heads_perbin.npz"), **out)
    print("heads_perbin.npz written")


def main_problosses():
    """losses_prob.npz: the soft-bin loss family of the reference (binDeltaLosses.py:109-208) with
    my_loss = geodesic_loss(reduce=False): values and gradients w.r.t. score and residual.
    Run with `python tests/golden/make_golden.py problosses`."""
    install_shims()
    import axisAngle as ref_aa
    import quaternion as ref_q
    import binDeltaLosses as ref_losses
    rng = np.random.default_rng(21)
    torch.manual_seed(21)
    B, K = 14, 6
    centers = rand_rotations(rng, K)[0]
    tmp = tempfile.mkdtemp()
    kfile = os.path.join(tmp, "k.pkl")
    with open(kfile, "wb") as f:
        pickle.dump(_PickleDict(centers), f)
    ydata, ydata_q = rand_rotations(rng, B)
    out = dict(centers=centers, ydata=ydata.astype(np.float32), ydata_q=ydata_q.astype(np.float32))
    score = torch.randn(B, K)
    res = 0.2 * torch.randn(B, 3)
    res_k = 0.2 * torch.randn(B, K, 3)
    res_q = 0.2 * torch.randn(B, 4)
    bins = torch.randint(0, K, (B,))
    prob = torch.softmax(torch.randn(B, K), 1)
    out.update(score=score.numpy(), res=res.numpy(), res_k=res_k.numpy(), res_q=res_q.numpy(),
               bins=bins.numpy(), prob=prob.numpy())
    yt = torch.from_numpy(out["ydata"])
    ytq = torch.from_numpy(out["ydata_q"])
    cases = {
        "prob": (ref_losses.ProbabilisticLoss(0.7, kfile, ref_aa.geodesic_loss(reduce=False)), res, bins, yt),
        "prob_multires": (ref_losses.ProbabilisticMultiresLoss(0.7, kfile, ref_aa.geodesic_loss(reduce=False)), res_k, bins, yt),
        "relaxed_q": (ref_losses.RelaXedProbabilisticLossQ(0.7, kfile, ref_q.geodesic_loss(reduce=False)), res_q, prob, ytq),
        "m3_geo": (ref_losses.loss_m3(0.7, kfile, ref_aa.geodesic_loss(reduce=False)), res, prob, yt),
    }
    for name, (crit, r, t0, t1) in cases.items():
        s_ = score.clone().requires_grad_(True)
        r_ = r.clone().requires_grad_(True)
        loss = crit([s_, r_], [t0, t1])
        loss.backward()
        out[name + "/loss"] = loss.detach().numpy()
        out[name + "/g_score"] = s_.grad.numpy()
        out[name + "/g_res"] = r_.grad.numpy()
    np.savez(os.path.join(OUT, "losses_prob.npz"), **out)
    print("losses_prob.npz written")


class _PickleDict:
    """Minimal stand-in for the pickled estimator the reference losses load: they only read
    `.cluster_centers_` and `.n_clusters` (binDeltaLosses.py:35-36, 138-139)."""

    def __init__(self, centers):
        self.cluster_centers_ = centers
        self.n_clusters = centers.shape[0]


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "perbin":
        main_perbin()
    elif len(sys.argv) > 1 and sys.argv[1] == "problosses":
        main_problosses()
    else:
        main()
