"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

A plain numpy / torch-fp32 restatement of the reference's arithmetic for the bin-and-delta pose hot
path (JHUVisionLab/multi-modal-regression), one function per reference call site, each citing the
file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module, and only as the checker / the timed CPU baseline.  The
product (multi-modal-regression_b200/) never imports it and has no CPU fallback.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against
outputs of the reference's OWN modules executed in the build container: tests/golden/make_golden.py
imports /root/reference/*.py (plus scikit-learn 1.9.0, the unpinned third-party dependency behind
kmeans.fit/predict) on seeded inputs and stores inputs + outputs in tests/golden/*.npz;
tests/test_oracle_golden.py checks every function below against those files.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

EPS = 1e-6   # helperFunctions.py:20


# --------------------------------------------------------------------------------------------------
# rotation helpers (numpy, per sample) — axisAngle.py:19-41, quaternion.py:18-29, 79-92
# --------------------------------------------------------------------------------------------------
def get_R(v):
    """axisAngle.py:33-41"""
    theta = np.linalg.norm(v)
    if theta < EPS:
        return np.eye(3)
    v = v / theta
    V = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + np.sin(theta) * V + (1 - np.cos(theta)) * np.dot(V, V)


def get_y(R):
    """axisAngle.py:19-29"""
    tR = 0.5 * (np.trace(R) - 1)
    theta = np.arccos(np.clip(tR, -1., 1.))
    tmp = 0.5 * (R - R.T)
    v = np.array([tmp[2, 1], tmp[0, 2], tmp[1, 0]])
    n = np.linalg.norm(v)
    v = v / n if n > EPS else np.zeros(3)
    return theta * v


def quat_get_y(R):
    """quaternion.py:18-29"""
    tR = 0.5 * (np.trace(R) - 1)
    theta = np.arccos(np.clip(tR, -1., 1.))
    tmp = 0.5 * (R - R.T)
    v = np.array([tmp[2, 1], tmp[0, 2], tmp[1, 0]])
    n = np.linalg.norm(v)
    if n > EPS:
        v = v / n
    else:
        theta, v = 0, np.zeros(3)
    return np.array([np.cos(theta / 2.), np.sin(theta / 2.) * v[0], np.sin(theta / 2.) * v[1],
                     np.sin(theta / 2.) * v[2]])


def convert_dictionary(aa_dict):
    """quaternion.py:79-92"""
    out = np.zeros((aa_dict.shape[0], 4))
    for i, x in enumerate(aa_dict):
        ang = np.linalg.norm(x)
        axis = x / ang if ang > EPS else np.zeros(3)
        y = np.array([np.cos(ang / 2.), np.sin(ang / 2.) * axis[0], np.sin(ang / 2.) * axis[1],
                      np.sin(ang / 2.) * axis[2]])
        out[i] = y / np.linalg.norm(y)
    return out


def rotation_matrix(az, el, ct):
    """helperFunctions.py:37-48"""
    ca, sa = np.cos(np.radians(az)), np.sin(np.radians(az))
    cb, sb = np.cos(np.radians(el)), np.sin(np.radians(el))
    cc, sc = np.cos(np.radians(ct)), np.sin(np.radians(ct))
    Ra = np.array([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1]])
    Rb = np.array([[1, 0, 0], [0, cb, -sb], [0, sb, cb]])
    Rc = np.array([[cc, -sc, 0], [sc, cc, 0], [0, 0, 1]])
    return np.dot(np.dot(Rc, Rb), Ra)


# --------------------------------------------------------------------------------------------------
# (d) evaluation metrics — axisAngle.py:45-95, quaternion.py:33-76
# --------------------------------------------------------------------------------------------------
def errors_aa(ygt, yhat):
    """per-sample part of axisAngle.get_error (axisAngle.py:48-60), degrees"""
    N = ygt.shape[0]
    err = np.zeros(N)
    for i in range(N):
        R = np.dot(get_R(ygt[i]).T, get_R(yhat[i]))
        theta = np.arccos(np.clip(0.5 * (np.trace(R) - 1), -1.0, 1.0))
        err[i] = np.rad2deg(np.abs(theta))
    return err


def errors_quat(ygt, yhat):
    """per-sample part of quaternion.get_error (quaternion.py:36-47), degrees"""
    N = ygt.shape[0]
    err = np.zeros(N)
    for i in range(N):
        tmp = np.clip(ygt[i, 0] * yhat[i, 0] + np.sum(ygt[i, 1:] * yhat[i, 1:]), -1.0, 1.0)
        err[i] = np.rad2deg(2.0 * np.arccos(np.abs(tmp)))
    return err


def get_error(ygt, yhat, quaternion=False):
    """(acc, medErr, errors) — axisAngle.py:61-66 / quaternion.py:48-51 (without the print)"""
    err = errors_quat(ygt, yhat) if quaternion else errors_aa(ygt, yhat)
    return 100 * np.sum(err < 30) / err.size, np.median(err), err


def get_error2(ygt, yhat, labels, num, quaternion=False):
    """axisAngle.py:86-95 / quaternion.py:69-76"""
    err = errors_quat(ygt, yhat) if quaternion else errors_aa(ygt, yhat)
    labels = np.squeeze(labels)
    med = np.zeros(num)
    for i in range(num):
        med[i] = np.median(err[labels == i])
    return np.mean(med)


# --------------------------------------------------------------------------------------------------
# (b) losses — torch fp32 with autograd, same op sequence as the reference
# --------------------------------------------------------------------------------------------------
def geodesic_loss_aa(ypred, ytrue, reduce=True):
    """axisAngle.py:110-120"""
    angle_pred = torch.norm(ypred, 2, 1)
    angle_true = torch.norm(ytrue, 2, 1)
    axis_pred = F.normalize(ypred)
    axis_true = F.normalize(ytrue)
    tmp = torch.abs(torch.cos(angle_true / 2) * torch.cos(angle_pred / 2) +
                    torch.sin(angle_true / 2) * torch.sin(angle_pred / 2) *
                    torch.sum(axis_true * axis_pred, dim=1))
    theta = 2.0 * torch.acos(torch.clamp(tmp, -1 + EPS, 1 - EPS))
    return torch.mean(theta) if reduce else theta


def geodesic_loss_quat(ypred, ytrue, reduce=True):
    """quaternion.py:156-163"""
    ypred = F.normalize(ypred)
    tmp = torch.abs(torch.sum(ytrue * ypred, dim=1))
    theta = 2.0 * torch.acos(torch.clamp(tmp, -1 + EPS, 1 - EPS))
    return torch.mean(theta) if reduce else theta


def rotmat_loss(ypred, ytrue):
    """RiemannianLoss.my_loss — binDeltaLosses.py:221-225"""
    tmp = torch.stack([torch.trace(torch.mm(ypred[i].t(), ytrue[i])) for i in range(ytrue.size(0))])
    return torch.mean(torch.acos(torch.clamp((tmp - 1.0) / 2, -1 + EPS, 1 - EPS)))


_PROJ = np.array([[0, 0, 0, 0, 0, -1, 0, 1, 0], [0, 0, 1, 0, 0, 0, -1, 0, 0],
                  [0, -1, 0, 1, 0, 0, 0, 0, 0]], dtype=np.float32)   # binDeltaLosses.py:216


def riemannian_terms(score, res, bin_true, R_true, key_poses):
    """(Lc, Lr) of RiemannianLoss.forward — binDeltaLosses.py:227-239"""
    l1 = F.cross_entropy(score, bin_true)
    ind = torch.max(score, dim=1)[1]
    angle = torch.norm(res, 2, 1)
    axis = F.normalize(res)
    axis = torch.mm(axis, torch.from_numpy(_PROJ)).view(-1, 3, 3)
    Id = torch.eye(3)
    y = torch.stack([Id + torch.sin(angle[i]) * axis[i] +
                     (1.0 - torch.cos(angle[i])) * torch.mm(axis[i], axis[i])
                     for i in range(angle.size(0))])
    y = torch.bmm(torch.index_select(key_poses, 0, ind), y)
    return l1, rotmat_loss(y, R_true)


def bin_delta_terms(score, res, bin_true, target, centers=None, pose='mse'):
    """(Lc, Lr) of SimpleLoss / GeodesicLoss / GeodesicLossQ — binDeltaLosses.py:22-28, 44-50, 66-72;
    script form learnGeodesicBDModel.py:175-179.  centers=None -> no key gather (SimpleLoss)."""
    l1 = F.cross_entropy(score, bin_true)
    y = res
    if centers is not None:
        ind = torch.max(score, dim=1)[1]
        y = torch.index_select(centers, 0, ind) + res
    if pose == 'mse':
        l2 = F.mse_loss(y, target)
    elif pose == 'aa':
        l2 = geodesic_loss_aa(y, target)
    elif pose == 'quat':
        l2 = geodesic_loss_quat(y, target)
    else:
        raise NameError(pose)
    return l1, l2


# --------------------------------------------------------------------------------------------------
# (c) assignment and k-means — binDeltaGenerators.py, learnKmeansDictionary.py:41-42 -> scikit-learn
# 1.9.0 (third party, unpinned by the reference): sklearn/cluster/_k_means_lloyd.pyx,
# _k_means_common.pyx, _kmeans.py::_kmeans_single_lloyd
# --------------------------------------------------------------------------------------------------
def e_step(X, centers, chunk=256):
    """sklearn _update_chunk_dense: labels = argmin_j (||c_j||^2 - 2 x.c_j), first minimum wins;
    X and centers float64; evaluated in 256-row chunks like the library."""
    X = np.asarray(X, dtype=np.float64)
    centers = np.asarray(centers, dtype=np.float64)
    cn = (centers ** 2).sum(1)
    labels = np.empty(X.shape[0], dtype=np.int32)
    for s in range(0, X.shape[0], chunk):
        d = cn[None, :] - 2.0 * X[s:s + chunk] @ centers.T
        labels[s:s + chunk] = np.argmin(d, axis=1)
    return labels


def predict_residual(ydata, centers):
    """GBDGenerator.__getitem__ — binDeltaGenerators.py:25-31 (predict input cast to float64, as the
    estimator requires; the residual is float64 - float64 -> .float())"""
    ydata = np.asarray(ydata)
    b = e_step(ydata.astype(np.float64), centers)
    res = (ydata - np.asarray(centers)[b, :]).astype(np.float32)
    return b.astype(np.int64), res


def riemannian_targets(ydata, centers):
    """RBDGenerator.__getitem__ — binDeltaGenerators.py:120, 129-138"""
    rotations_dict = np.stack([get_R(centers[i]) for i in range(centers.shape[0])])
    ydata = np.asarray(ydata)
    rot = np.stack([get_R(ydata[i]) for i in range(ydata.shape[0])])
    b = e_step(ydata.astype(np.float64), centers)
    res = np.stack([get_y(np.dot(rotations_dict[b[i]].T, rot[i])) for i in range(ydata.shape[0])])
    return b.astype(np.int64), res.astype(np.float32), rot.astype(np.float32)


def quatdot_assign(q, keys):
    """learnObjectnetModel.py:108-109 (per sample)"""
    q = np.asarray(q)
    bins = np.zeros(q.shape[0], dtype=np.int64)
    res = np.zeros((q.shape[0], 4), dtype=np.float32)
    for i in range(q.shape[0]):
        b = np.argmax(np.abs(np.dot(keys, q[i])))
        bins[i] = b
        res[i] = q[i] - keys[b, :]
    return bins, res


def soft_assign(ydata, centers, gamma=10.0):
    """XPBDGeneratorQ.__getitem__ — binDeltaGenerators.py:104-108"""
    from scipy.spatial.distance import cdist
    p = np.exp(-gamma * cdist(ydata, centers, 'sqeuclidean'))
    p = p / np.sum(p, axis=1, keepdims=True)
    return p.astype(np.float32), (ydata - np.dot(p, centers)).astype(np.float32)


def kmeans_lloyd(X, init, max_iter=300, tol=1e-4):
    """KMeans(K, init=<array>, n_init=1, algorithm='lloyd').fit(X) restated from scikit-learn 1.9.0:
    _kmeans.py:1488-1546 (mean-centring), _tolerance (285), _kmeans_single_lloyd (689-758),
    lloyd_iter_chunked_dense + _relocate_empty_clusters_dense + _average_centers + _center_shift."""
    X = np.array(X, dtype=np.float64)
    mean = X.mean(axis=0)
    X -= mean
    centers = np.array(init, dtype=np.float64) - mean
    tol_abs = np.mean(np.var(X, axis=0)) * tol
    K = centers.shape[0]
    labels_old = np.full(X.shape[0], -1, dtype=np.int32)
    strict = False
    n_iter = 0
    for i in range(max_iter):
        labels = e_step(X, centers)
        sums = np.zeros_like(centers)
        np.add.at(sums, labels, X)
        w = np.bincount(labels, minlength=K).astype(np.float64)
        empty = np.where(w == 0)[0]
        if empty.size:
            dist = ((X - centers[labels]) ** 2).sum(axis=1)
            far = np.argpartition(dist, -empty.size)[:-empty.size - 1:-1]
            if np.max(dist) != 0:
                for idx in range(empty.size):
                    new_c, f = empty[idx], far[idx]
                    old_c = labels[f]
                    sums[old_c] -= X[f]
                    sums[new_c] = X[f]
                    w[new_c] = 1
                    w[old_c] -= 1
        big = int(np.argmax(w))
        new = np.empty_like(centers)
        for j in range(K):          # _average_centers (in-order, see _k_means_common.pyx:286-295)
            if w[j] > 0:
                sums[j] *= 1.0 / w[j]
            else:
                sums[j] = sums[big]
            new[j] = sums[j]
        shift_tot = ((new - centers) ** 2).sum()
        centers = new
        n_iter = i + 1
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if shift_tot <= tol_abs:
            break
        labels_old = labels
    if not strict:
        labels = e_step(X, centers)
    inertia = ((X - centers[labels]) ** 2).sum()
    return dict(centers=centers + mean, labels=labels, inertia=float(inertia), n_iter=n_iter)


# --------------------------------------------------------------------------------------------------
# (a) heads — binDeltaModels.py:62-121, learnJointCatPoseModel_weighted.py:107-115,
# objectnetHelperFunctions.py:110-172 (torch fp32 modules with the reference's layer names)
# --------------------------------------------------------------------------------------------------
class Mlp3(nn.Module):
    """bin_3layer / res_3layer / poseModels.model_3layer — binDeltaModels.py:62-91"""

    def __init__(self, N0, N1, N2, Nout):
        super().__init__()
        self.fc1 = nn.Linear(N0, N1, bias=False)
        self.bn1 = nn.BatchNorm1d(N1)
        self.fc2 = nn.Linear(N1, N2, bias=False)
        self.bn2 = nn.BatchNorm1d(N2)
        self.fc3 = nn.Linear(N2, Nout)

    def forward(self, x):
        x = F.relu(self.bn1(self.fc1(x)))
        x = F.relu(self.bn2(self.fc2(x)))
        return self.fc3(x)


class OneBinDeltaHeads(nn.Module):
    """The head part of OneBinDeltaModel (binDeltaModels.py:109-120) on a given feature tensor, with
    either a class label (one-hot mixing, 116-119) or soft mixing weights (joint model,
    learnJointCatPoseModel_weighted.py:110-115)."""

    def __init__(self, num_classes, num_clusters, N0, N1, N2, ndim):
        super().__init__()
        self.num_classes = num_classes
        self.bin_models = nn.ModuleList([Mlp3(N0, N1, N2, num_clusters) for _ in range(num_classes)])
        self.res_models = nn.ModuleList([Mlp3(N0, N1, N2, ndim) for _ in range(num_classes)])

    def forward(self, x, label=None, mix=None):
        y1 = torch.stack([m(x) for m in self.bin_models]).permute(1, 2, 0)
        y2 = torch.stack([m(x) for m in self.res_models]).permute(1, 2, 0)
        if mix is None:
            mix = torch.zeros(label.size(0), self.num_classes).scatter_(1, label, 1.0)
        mix = mix.to(y1.dtype).unsqueeze(2)
        return [torch.squeeze(torch.bmm(y1, mix), 2), torch.squeeze(torch.bmm(y2, mix), 2)]


class ObjectnetHeads(nn.Module):
    """The head part of objectnetHelperFunctions.OneBinDeltaModel (155-172): one 3-layer MLP pair on
    cat(features, onehot(label))."""

    def __init__(self, num_classes, dict_size=200, n0=2048, n1=1000, n2=500, dim=3):
        super().__init__()
        self.num_classes = num_classes
        self.bin_model = Mlp3(n0 + num_classes, n1, n2, dict_size)
        self.res_model = Mlp3(n0 + num_classes, n1, n2, dim)

    def forward(self, x, label):
        onehot = torch.zeros(label.size(0), self.num_classes).scatter_(1, label, 1.0)
        x = torch.cat((x, onehot), dim=1)
        return [self.bin_model(x), self.res_model(x)]


class Mlp2(nn.Module):
    """bin_2layer / res_2layer — binDeltaModels.py:36-59"""

    def __init__(self, N0, N1, Nout):
        super().__init__()
        self.fc1 = nn.Linear(N0, N1, bias=False)
        self.bn1 = nn.BatchNorm1d(N1)
        self.fc2 = nn.Linear(N1, Nout)

    def forward(self, x):
        return self.fc2(F.relu(self.bn1(self.fc1(x))))


class OneDeltaPerBinHeads(nn.Module):
    """The head part of OneDeltaPerBinModel / ProbabilisticOneDeltaPerBinModel
    (binDeltaModels.py:124-178): C bin heads + C*K two-layer delta heads; the delta is picked by the
    class label and then by the argmax bin (146-149), or all K deltas of the class are returned
    (probabilistic: 174-176)."""

    def __init__(self, num_classes, num_clusters, N0, N1, N2, N3, ndim):
        super().__init__()
        self.num_classes, self.num_clusters, self.ndim = num_classes, num_clusters, ndim
        self.bin_models = nn.ModuleList([Mlp3(N0, N1, N2, num_clusters) for _ in range(num_classes)])
        self.res_models = nn.ModuleList([Mlp2(N0, N3, ndim) for _ in range(num_classes * num_clusters)])

    def forward(self, x, class_label, probabilistic=False):
        y1 = torch.stack([m(x) for m in self.bin_models]).permute(1, 2, 0)
        y2 = torch.stack([m(x) for m in self.res_models])
        y2 = y2.view(self.num_classes, self.num_clusters, -1, self.ndim).permute(1, 2, 3, 0)
        cl = torch.zeros(class_label.size(0), self.num_classes).scatter_(1, class_label, 1.0).unsqueeze(2)
        y1 = torch.squeeze(torch.bmm(y1, cl), 2)
        y2 = torch.squeeze(torch.matmul(y2, cl), 3)                     # [K, B, ndim]
        if probabilistic:
            return [y1, y2.permute(1, 0, 2)]
        _, pose_label = torch.max(y1, dim=1, keepdim=True)
        pl = torch.zeros(pose_label.size(0), self.num_clusters).scatter_(1, pose_label, 1.0).unsqueeze(2)
        return [y1, torch.squeeze(torch.bmm(y2.permute(1, 2, 0), pl), 2)]


# --------------------------------------------------------------------------------------------------
# round 2: loss_m2, get_gamma, test-time compositions, mySGD, get_accuracy, detection-metric helpers
# --------------------------------------------------------------------------------------------------
def loss_m2(score, res, bin_true, res_true, alpha):
    """binDeltaLosses.py:280-297: CE + alpha * MSE(residual, res_true[b, argmax score_b, :]) — the
    one-hot bmm over ytrue[1].permute(0, 2, 1) picks the row of the predicted bin."""
    l1 = F.cross_entropy(score, bin_true)
    ind = torch.argmax(score, dim=1)
    yres = res_true[torch.arange(score.shape[0]), ind]
    return l1 + alpha * F.mse_loss(res, yres)


def get_gamma(kmeans_dict):
    """helperFunctions.py:51-58"""
    c = np.asarray(kmeans_dict, dtype=np.float64)
    D = ((c[:, None, :] - c[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(D, np.inf)
    return 1.0 / (2.0 * D.min())


def compose_add(score, res, dictionary):
    """learnGeodesicBDModel.py:217-219"""
    return np.asarray(dictionary)[np.argmax(score, axis=1), :] + res


def compose_normalize(score, res, dictionary):
    """learnGeodesicBDModel_quaternion.py:217-218"""
    y = compose_add(score, res, dictionary)
    return y / np.maximum(np.linalg.norm(y, 2, 1, True), 1e-10)


def compose_riemannian(score, res, rot_dict):
    """learnRiemannianBDModel.py:247"""
    b = np.argmax(score, axis=1)
    return np.stack([get_y(np.dot(rot_dict[b[j]], get_R(res[j]))) for j in range(b.shape[0])])


def cyclic_step_size(step, c, alpha1, alpha2):
    """helperFunctions.py:111-116"""
    t = (np.fmod(step - 1, c) + 1) / c
    if t <= 0.5:
        return (1 - 2 * t) * alpha1 + 2 * t * alpha2
    return 2 * (1 - t) * alpha2 + (2 * t - 1) * alpha1


def my_sgd(params, grad_fn, n_steps, c, alpha1, alpha2, momentum=0.0, dampening=0.0, weight_decay=0.0,
           nesterov=False):
    """helperFunctions.mySGD (62-120) on numpy arrays: params list, grad_fn(params) -> grads list.
    Returns the flattened parameters after every step."""
    params = [np.array(p, dtype=np.float32) for p in params]
    bufs = [None] * len(params)
    traj = []
    for step in range(1, n_steps + 1):
        grads = grad_fn(params)
        lr = np.float32(cyclic_step_size(step, c, alpha1, alpha2))
        for i, (p, g) in enumerate(zip(params, grads)):
            d_p = np.array(g, dtype=np.float32)
            if weight_decay != 0:
                d_p = d_p + np.float32(weight_decay) * p
            if momentum != 0:
                if bufs[i] is None:
                    bufs[i] = d_p.copy()
                else:
                    bufs[i] = np.float32(momentum) * bufs[i] + np.float32(1 - dampening) * d_p
                d_p = d_p + np.float32(momentum) * bufs[i] if nesterov else bufs[i]
            params[i] = p - lr * d_p
        traj.append(np.concatenate([p.ravel() for p in params]))
    return np.stack(traj)


def get_accuracy(ytrue, ypred, num_classes):
    """helperFunctions.py:123-130"""
    return float(np.mean([np.sum((ytrue == i) * (ypred == i)) / np.sum(ytrue == i) for i in range(num_classes)]))
