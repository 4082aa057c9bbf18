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


def _fake_dataset(root, classes, counts, rng, png=True):
    """A synthetic dataset directory in the reference's on-disk layout (dataGenerators.py:34-37,
    53): <class>_info.mat with `image_names`, <class>/<name>.png.  Names carry the Euler angles the
    way parse_name reads them (helperFunctions.py:24-33): syn_model_a<az>_e<el>_t<ct>_d<dist>."""
    import scipy.io as spio
    from PIL import Image
    names_all = []
    for c, n in zip(classes, counts):
        names = []
        for j in range(n):
            az, el, ct = rng.uniform(0, 360), rng.uniform(-60, 60), rng.uniform(-40, 40)
            if j == 0:
                az, el, ct = 0.0, 0.0, 0.0                       # identity rotation: get_y -> 0
            names.append("%s_m%02d_a%.4f_e%.4f_t%.4f_d%.3f" % (c[:3], j, az, el, ct, rng.uniform(1, 3)))
        spio.savemat(os.path.join(root, c + "_info.mat"), {"image_names": np.array(names, dtype=object)})
        os.makedirs(os.path.join(root, c), exist_ok=True)
        if png:
            for nme in names:
                Image.fromarray(rng.integers(0, 255, (8, 8, 3), dtype=np.uint8)).save(
                    os.path.join(root, c, nme + ".png"))
        names_all.append(names)
    return names_all


def main_generators():
    """generators.npz: the reference's Dataset classes END TO END on a synthetic dataset directory —
    dataGenerators.ImagesAll (names -> Euler -> R -> pose target) under binDeltaGenerators.GBDGenerator /
    GBDGeneratorQ / XPBDGeneratorQ / RBDGenerator (real and render), and objectnetHelperFunctions.
    TrainImages / TestImages.  Run with `python tests/golden/make_golden.py generators`."""
    install_shims()
    import helperFunctions as ref_h
    import binDeltaGenerators as ref_gen
    import objectnetHelperFunctions as ref_on
    from sklearn.cluster import KMeans as SKK
    rng = np.random.default_rng(31)
    orig_predict = SKK.predict

    def predict64(self, X):
        X = np.asarray(X, dtype=np.float64)
        self.n_features_in_ = X.shape[1]
        return orig_predict(self, X)

    SKK.predict = predict64
    root = tempfile.mkdtemp()
    classes = list(ref_h.classes)
    counts = [5, 7, 6, 5, 7, 6, 5, 7, 6, 5, 7, 4]
    names = _fake_dataset(root, classes, counts, rng)
    Kd = 16
    train, _ = rand_rotations(rng, 3000)
    sk = SKK(n_clusters=Kd, init=train[:Kd].copy(), n_init=1, max_iter=300, tol=1e-4,
             algorithm="lloyd").fit(train)
    os.makedirs(os.path.join(root, "data"), exist_ok=True)
    dfile = os.path.join(root, "data", "kmeans_dictionary_axis_angle_16.pkl")
    with open(dfile, "wb") as f:
        pickle.dump(sk, f)
    out = dict(centers=sk.cluster_centers_, counts=np.array(counts), classes=np.array(classes),
               names=np.array([n for per in names for n in per]))
    n_items = max(counts)
    for cls, key, kinds in ((ref_gen.GBDGenerator, "gbd", ("real", "render")),
                            (ref_gen.GBDGeneratorQ, "gbdq", ("real",)),
                            (ref_gen.XPBDGeneratorQ, "xpbdq", ("render",)),
                            (ref_gen.RBDGenerator, "rbd", ("real",))):
        for kind in kinds:
            with open(dfile, "wb") as f:          # GBDGeneratorQ mutates the estimator it loads
                pickle.dump(sk, f)
            g = cls(root, kind, dfile)
            assert len(g) == n_items
            items = [g[i] for i in range(n_items)]
            k = "%s_%s" % (key, kind)
            out[k + "_ydata"] = np.stack([s["ydata"].numpy() for s in items])
            out[k + "_label"] = np.stack([s["label"].numpy() for s in items])
            out[k + "_bin"] = np.stack([s["ydata_bin"].numpy() for s in items])
            out[k + "_res"] = np.stack([s["ydata_res"].numpy() for s in items])
            out[k + "_xmean"] = np.stack([s["xdata"].mean(dim=(1, 2, 3)).numpy() for s in items])
            if "ydata_rot" in items[0]:
                out[k + "_rot"] = np.stack([s["ydata_rot"].numpy() for s in items])
    # ObjectNet datasets (objectnetHelperFunctions.py:23-107): the dictionary path is relative
    with open(dfile, "wb") as f:
        pickle.dump(sk, f)
    cwd = os.getcwd()
    os.chdir(root)
    try:
        tr = ref_on.TrainImages(root, classes[:5], dict_size=16)
        items = [tr[i] for i in range(len(tr))]
        for fld in ("ydata", "label", "ydata_bin", "ydata_res"):
            out["on_train_" + fld] = np.stack([s[fld].numpy() for s in items])
        # numpy >= 2 refuses `0-d ndarray * Tensor` (objectnetHelperFunctions.py:102 was written for
        # numpy 1.x, where it gave a [1] long tensor): np.squeeze of the one-element prediction is
        # handed over as a python int, which indexes and multiplies the same way
        class _NP:
            def __getattr__(self, k):
                return getattr(np, k)

            @staticmethod
            def squeeze(a):
                return int(np.squeeze(a))
        ref_on.np = _NP()
        te = ref_on.TestImages(root, classes[:5], dict_size=16)
        items = [te[i] for i in range(len(te))]
        for fld in ("ydata", "label", "ydata_bin", "ydata_res"):
            out["on_test_" + fld] = np.stack([s[fld].numpy() for s in items])
    finally:
        os.chdir(cwd)
    np.savez(os.path.join(OUT, "generators.npz"), **out)
    print("generators.npz written")


def _reference_class(path, name, namespace):
    """Compile ONE class definition of a reference script (the scripts run argparse + training at
    import, so they cannot be imported) in `namespace`; nothing is written to disk."""
    import ast
    with open(path) as f:
        tree = ast.parse(f.read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name)
    mod = ast.Module(body=[node], type_ignores=[])
    exec(compile(mod, path, "exec"), namespace)
    return namespace[name]


def main_joint():
    """joint.npz: the reference's JointCatPoseModel (learnJointCatPoseModel_weighted.py:94-126),
    compiled from the script's own class statement, re-parenting a reference OneBinDeltaModel
    (multires False) and OneDeltaPerBinModel (multires True): forward in train mode + gradients of a
    seeded linear functional, BatchNorm statistics, eval forward.
    Run with `python tests/golden/make_golden.py joint`."""
    install_shims()
    import torch.nn.functional as F
    from torch.autograd import Variable
    import binDeltaModels as ref_models
    ref_models.resnet_model = lambda *a, **k: nn.Identity()
    Cc, Kc, N0, N1, N2, N3, nd, Bh = 3, 4, 64, 40, 24, 12, 3, 10
    out = dict(dims=np.array([Cc, Kc, N0, N1, N2, N3, nd, Bh]))
    for multires in (False, True):
        torch.manual_seed(41 + int(multires))
        tag = "mr/" if multires else "sr/"
        args = types.SimpleNamespace(multires=multires)
        ns = dict(torch=torch, nn=nn, F=F, Variable=Variable, args=args, N0=N0, num_classes=Cc)
        cls = _reference_class(os.path.join(REF, "learnJointCatPoseModel_weighted.py"),
                               "JointCatPoseModel", ns)
        if multires:
            base = ref_models.OneDeltaPerBinModel("resnet", Cc, Kc, N0, N1, N2, N3, nd)
        else:
            base = ref_models.OneBinDeltaModel("resnet", Cc, Kc, N0, N1, N2, nd)
        for m in base.modules():
            if isinstance(m, nn.BatchNorm1d):
                m.weight.data.uniform_(0.5, 1.5)
                m.bias.data.uniform_(-0.3, 0.3)
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 2.0)
        model = cls(base)
        for k, v in model.state_dict().items():
            out[tag + "sd0/" + k] = v.clone().numpy()
        x = torch.randn(Bh, N0)
        out[tag + "x"] = x.numpy()
        model.train()
        xr = x.clone().requires_grad_(True)
        y0, y1, y2 = model(xr)
        ws = [torch.randn_like(y) for y in (y0, y1, y2)]
        sum((y * w).sum() for y, w in zip((y0, y1, y2), ws)).backward()
        for i, (y, w) in enumerate(zip((y0, y1, y2), ws)):
            out[tag + "train_y%d" % i] = y.detach().numpy()
            out[tag + "w%d" % i] = w.numpy()
        out[tag + "train_gx"] = xr.grad.numpy()
        for k, p in model.named_parameters():
            out[tag + "train_grad/" + k] = p.grad.numpy()
        for k, v in model.state_dict().items():
            if "running" in k or "num_batches" in k:
                out[tag + "train_sd/" + k] = v.clone().numpy()
        model.eval()
        with torch.no_grad():
            e = model(x)
        for i, y in enumerate(e):
            out[tag + "eval_y%d" % i] = y.numpy()
    np.savez(os.path.join(OUT, "joint.npz"), **out)
    print("joint.npz written")


def main_objectnet():
    """objectnet_head.npz: objectnetHelperFunctions.OneBinDeltaModel (155-172) of the reference on
    small layer sizes: train forward + gradients, BatchNorm statistics, eval forward.
    Run with `python tests/golden/make_golden.py objectnet`."""
    install_shims()
    import objectnetHelperFunctions as ref_on
    ref_on.resnet_model = lambda *a, **k: nn.Identity()
    torch.manual_seed(51)
    C, K, n0, n1, n2, nd, B = 5, 16, 59, 40, 24, 3, 12          # n0 + C = 64 input columns
    model = ref_on.OneBinDeltaModel(C, K, n0, n1, n2, nd)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm1d):
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.uniform_(-0.3, 0.3)
            m.running_mean.uniform_(-0.2, 0.2)
            m.running_var.uniform_(0.5, 2.0)
    out = dict(dims=np.array([C, K, n0, n1, n2, nd, B]))
    for k, v in model.state_dict().items():
        out["sd0/" + k] = v.clone().numpy()
    x = torch.randn(B, n0)
    label = torch.randint(0, C, (B, 1))
    out.update(x=x.numpy(), label=label.numpy())
    model.train()
    xr = x.clone().requires_grad_(True)
    y1, y2 = model(xr, label)
    w1, w2 = torch.randn_like(y1), torch.randn_like(y2)
    ((y1 * w1).sum() + (y2 * w2).sum()).backward()
    out.update(train_y1=y1.detach().numpy(), train_y2=y2.detach().numpy(), w1=w1.numpy(), w2=w2.numpy(),
               train_gx=xr.grad.numpy())
    for k, p in model.named_parameters():
        out["train_grad/" + k] = p.grad.numpy()
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["train_sd/" + k] = v.clone().numpy()
    model.eval()
    with torch.no_grad():
        e1, e2 = model(x, label)
    out.update(eval_y1=e1.numpy(), eval_y2=e2.numpy())
    np.savez(os.path.join(OUT, "objectnet_head.npz"), **out)
    print("objectnet_head.npz written")


def main_misc():
    """misc_r2.npz: loss_m2 (binDeltaLosses.py:280-297), get_gamma (helperFunctions.py:51-58), the
    prediction compositions of the scripts' testing() loops (learnGeodesicBDModel.py:217-219,
    learnGeodesicBDModel_quaternion.py:217-218, learnRiemannianBDModel.py:247 — the script lines are
    evaluated here with the reference's get_y / get_R), the mySGD cyclical-LR optimizer
    (helperFunctions.py:62-120) and get_accuracy (123-130).
    Run with `python tests/golden/make_golden.py misc`."""
    install_shims()
    import axisAngle as ref_aa
    import binDeltaLosses as ref_losses
    import helperFunctions as ref_h
    rng = np.random.default_rng(61)
    torch.manual_seed(61)
    out = {}
    # loss_m2: per-bin residual targets [B, K, ndim], the row of the ARGMAX bin is regressed
    B, K, nd = 9, 7, 3
    score = torch.randn(B, K)
    score[2, 1] = score[2, 4] = score[2].max() + 1.0               # tie -> lowest index
    res = 0.3 * torch.randn(B, nd)
    bins = torch.randint(0, K, (B,))
    res_true = 0.3 * torch.randn(B, K, nd)
    s_ = score.clone().requires_grad_(True)
    r_ = res.clone().requires_grad_(True)
    loss = ref_losses.loss_m2(0.6, K)([s_, r_], [bins, res_true])
    loss.backward()
    out.update(m2_score=score.numpy(), m2_res=res.numpy(), m2_bins=bins.numpy(),
               m2_res_true=res_true.numpy(), m2_alpha=0.6, m2_loss=loss.detach().numpy(),
               m2_g_score=s_.grad.numpy(), m2_g_res=r_.grad.numpy())
    # get_gamma
    cen = rand_rotations(rng, 40)[0]
    out.update(gamma_centers=cen, gamma=ref_h.get_gamma(cen))
    # testing() compositions
    N, Kt = 50, 12
    kdict = rand_rotations(rng, Kt)[0]
    sc = rng.standard_normal((N, Kt)).astype(np.float32)
    sc[5, 2] = sc[5, 9] = sc[5].max() + 1
    rs = (0.3 * rng.standard_normal((N, 3))).astype(np.float32)
    rs[7] = 0.0
    ybin = np.argmax(sc, axis=1)
    out.update(t_dict=kdict, t_score=sc, t_res=rs, t_add=kdict[ybin, :] + rs)
    qdict = rand_rotations(rng, Kt)[1]
    rs4 = (0.3 * rng.standard_normal((N, 4))).astype(np.float32)
    y = qdict[ybin, :] + rs4
    out.update(t_qdict=qdict, t_res4=rs4, t_quat=y / np.maximum(np.linalg.norm(y, 2, 1, True), 1e-10))
    rot_dict = np.stack([ref_aa.get_R(kdict[i]) for i in range(Kt)])
    out.update(t_rotdict=rot_dict, t_riem=np.stack(
        [ref_aa.get_y(np.dot(rot_dict[ybin[j]], ref_aa.get_R(rs[j]))) for j in range(N)]))
    # mySGD: 7 steps of the cyclical learning rate on a two-tensor problem (c = 4)
    import warnings
    warnings.simplefilter("ignore")
    p1 = torch.randn(5, 3).requires_grad_(True)
    p2 = torch.randn(4).requires_grad_(True)
    out.update(sgd_p1=p1.detach().numpy().copy(), sgd_p2=p2.detach().numpy().copy())
    tgt1, tgt2 = torch.randn(5, 3), torch.randn(4)
    out.update(sgd_t1=tgt1.numpy(), sgd_t2=tgt2.numpy())
    for name, kw in (("plain", {}), ("mom", dict(momentum=0.9, weight_decay=1e-2)),
                     ("nest", dict(momentum=0.8, nesterov=True, dampening=0.0))):
        q1 = p1.detach().clone().requires_grad_(True)
        q2 = p2.detach().clone().requires_grad_(True)
        opt = ref_h.mySGD([q1, q2], c=4, alpha1=1e-1, alpha2=1e-3, **kw)
        traj = []
        for it in range(7):
            opt.zero_grad()
            (((q1 - tgt1) ** 2).sum() + ((q2 - tgt2) ** 4).sum()).backward()
            opt.step()
            traj.append(np.concatenate([q1.detach().numpy().ravel(), q2.detach().numpy().ravel()]))
        out["sgd_traj_" + name] = np.stack(traj)
    yt = rng.integers(0, 5, 200)
    yp = np.where(rng.uniform(size=200) < 0.7, yt, rng.integers(0, 5, 200))
    out.update(acc_true=yt, acc_pred=yp, acc=ref_h.get_accuracy(yt, yp, 5))
    np.savez(os.path.join(OUT, "misc_r2.npz"), **out)
    print("misc_r2.npz written")


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
    elif len(sys.argv) > 1 and sys.argv[1] in ("generators", "joint", "objectnet", "misc"):
        {"generators": main_generators, "joint": main_joint, "objectnet": main_objectnet,
         "misc": main_misc}[sys.argv[1]]()
    else:
        main()
