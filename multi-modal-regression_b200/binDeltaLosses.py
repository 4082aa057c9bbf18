"""Drop-in for the reference's binDeltaLosses module (binDeltaLosses.py:1-334).

Each composite loss `crit(ypred=[score, residual], ytrue=[bin, target])` is ONE fused launch of
bdp_bd_loss_fwd_bwd (cross-entropy over the bins, argmax key gather, pose composition, pose loss
and both gradients); autograd only rescales the stored gradients by the upstream scalars.  Every
class also exposes `.terms(ypred, ytrue) -> (Lc, Lr)` for scripts that weight the two terms
themselves with a python float (learnGeodesicBDModel.py:180, 185).

The soft-bin families (RelaXed*/Probabilistic*/loss_m2..m4) keep the reference's
call signature; their K-wide expectation is evaluated with per-sample launches of the same pose
kernels (reduce=False) instead of the reference's python loop over clusters.
"""
import pickle

import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from bdpose import ops
from bdpose import _lib as L
import axisAngle as _aa
import quaternion as _quat


def _load_centers(kmeans_file):
    with open(kmeans_file, 'rb') as f:
        km = pickle.load(f)
    return km, np.asarray(km.cluster_centers_)


def _pose_mode_of(my_loss, ndim):
    """Map the reference's `my_loss` argument onto a fused pose mode (None -> MSE)."""
    if my_loss is None or isinstance(my_loss, nn.MSELoss):
        return L.POSE_MSE
    if isinstance(my_loss, _aa.geodesic_loss):
        return L.POSE_GEODESIC_AA
    if isinstance(my_loss, _quat.geodesic_loss):
        return L.POSE_GEODESIC_Q
    return None   # foreign callable: composed un-fused below


class _FusedBinDelta(nn.Module):
    """Shared machinery: Lc + alpha * Lr with (Lc, Lr) from the fused kernel."""

    def __init__(self, alpha, pose_mode, keys=None, my_loss=None):
        super().__init__()
        self.alpha = alpha
        self.pose_mode = pose_mode
        self.my_loss = my_loss
        if keys is not None:
            self.register_buffer('cluster_centers_', torch.as_tensor(keys).float().cuda(),
                                 persistent=False)
        else:
            self.cluster_centers_ = None

    def terms(self, ypred, ytrue):
        score, res = ypred[0], ypred[1]
        use_keys = self.cluster_centers_ is not None
        if self.pose_mode is None:
            # a user-supplied pose loss: CE + argmax from the fused kernel, composition in torch
            lc, _, ind = ops.bd_loss(score, ytrue[0], None, None, None, L.POSE_NONE, False)
            y = res + self.cluster_centers_.index_select(0, ind) if use_keys else res
            return lc, self.my_loss(y, ytrue[1])
        out = ops.bd_loss(score, ytrue[0], res, ytrue[1], self.cluster_centers_, self.pose_mode,
                          use_keys)
        return out[0], out[1]

    def forward(self, ypred, ytrue):
        lc, lr = self.terms(ypred, ytrue)
        return lc + self.alpha * lr


class SimpleLoss(_FusedBinDelta):
    """CE(bin) + alpha * MSE(residual, ydata_res) — binDeltaLosses.py:16-28"""

    def __init__(self, alpha):
        super().__init__(alpha, L.POSE_MSE)


class loss_m0(SimpleLoss):
    """binDeltaLosses.py:243-256 (same arithmetic as SimpleLoss)"""


class GeodesicLoss(_FusedBinDelta):
    """CE(bin) + alpha * my_loss(centers[argmax score] + residual, ydata) — binDeltaLosses.py:31-50"""

    def __init__(self, alpha, kmeans_file, my_loss=None):
        _, centers = _load_centers(kmeans_file)
        super().__init__(alpha, _pose_mode_of(my_loss, 3), centers, my_loss)


class loss_m1(GeodesicLoss):
    """binDeltaLosses.py:259-277 (same arithmetic as GeodesicLoss)"""


class GeodesicLossQ(_FusedBinDelta):
    """As GeodesicLoss with the dictionary converted to unit quaternions — binDeltaLosses.py:53-72"""

    def __init__(self, alpha, kmeans_file, my_loss=None):
        _, centers = _load_centers(kmeans_file)
        super().__init__(alpha, _pose_mode_of(my_loss, 4), _quat.convert_dictionary(centers), my_loss)


class RiemannianLoss(nn.Module):
    """CE(bin) + alpha * mean acos(clamp((tr((K[argmax] exp([r]x))^T R) - 1)/2)) —
    binDeltaLosses.py:211-239.  pose_dict: [K,3,3] key rotations (numpy)."""

    def __init__(self, alpha, pose_dict):
        super().__init__()
        self.alpha = alpha
        self.register_buffer('key_poses', torch.as_tensor(np.asarray(pose_dict)).float().cuda(),
                             persistent=False)

    def my_loss(self, ypred, ytrue):
        """mean geodesic angle between predicted and true rotation MATRICES [B,3,3]
        (binDeltaLosses.py:221-225)."""
        return ops.pose_loss(ypred.reshape(-1, 9), ytrue.reshape(-1, 9), L.POSE_ROTMAT)

    def terms(self, ypred, ytrue):
        out = ops.bd_loss(ypred[0], ytrue[0], ypred[1], ytrue[1].reshape(-1, 9),
                          self.key_poses.reshape(-1, 9), L.POSE_RIEMANNIAN, True)
        return out[0], out[1]

    def forward(self, ypred, ytrue):
        lc, lr = self.terms(ypred, ytrue)
        return lc + self.alpha * lr


# ---- soft-bin families ------------------------------------------------------------
# True: the expectation over the bins runs as one fused launch (bdp_expected_pose_loss); False: the
# reference's loop over the K bins on the per-row pose-loss kernel (kept for foreign my_loss callables)
FUSE_EXPECTED_POSE = True


def _kl(logits, soft_bins):
    # nn.KLDivLoss() default reduction ('mean' over all elements), as the reference constructs it
    return F.kl_div(F.log_softmax(logits, dim=1), soft_bins, reduction='mean')


def _rows(my_loss, pred, target):
    """Per-sample pose loss [B] for a reduce=False criterion."""
    if my_loss is None or isinstance(my_loss, nn.MSELoss):
        return ((pred - target) ** 2)     # nn.MSELoss(reduce=False): elementwise, as loss_m3 uses it
    return my_loss(pred, target)


def _expected_pose_loss(score, my_loss, target, poses_per_bin):
    """mean_b sum_k softmax(score)_bk * loss(target_b, pose_bk); poses_per_bin(k) -> [B, ndim]."""
    K = score.shape[1]
    cols = [_rows(my_loss, target, poses_per_bin(k)) for k in range(K)]
    l2 = torch.stack(cols)                                   # [K, B] (or [K, B, ndim] for MSE)
    if l2.dim() == 2:
        return torch.mean(torch.sum(F.softmax(score, dim=1) * l2.t(), dim=1))
    return torch.mean(torch.sum(F.softmax(score, dim=1) * torch.t(l2), dim=1))


class SimpleRelaXedLoss(nn.Module):
    """KL(soft bins) + alpha * MSE(residual) — binDeltaLosses.py:75-88"""

    def __init__(self, alpha):
        super().__init__()
        self.alpha = alpha

    def forward(self, ypred, ytrue):
        lr = ops.pose_loss(ypred[1], ytrue[1], L.POSE_MSE)
        return _kl(ypred[0], ytrue[0]) + self.alpha * lr


class RelaXedLoss(nn.Module):
    """KL(soft bins) + alpha * my_loss(centers[argmax] + residual, ydata) — binDeltaLosses.py:91-108"""

    def __init__(self, alpha, kmeans_file, my_loss):
        super().__init__()
        self.alpha = alpha
        _, centers = _load_centers(kmeans_file)
        self.register_buffer('cluster_centers_', torch.as_tensor(centers).float().cuda(),
                             persistent=False)
        self.my_loss = my_loss

    def forward(self, ypred, ytrue):
        ind = torch.argmax(ypred[0], dim=1)
        y = self.cluster_centers_.index_select(0, ind) + ypred[1]
        return _kl(ypred[0], ytrue[0]) + self.alpha * self.my_loss(y, ytrue[1])


class _ProbBase(nn.Module):
    use_kl = True
    per_bin_delta = False

    def __init__(self, alpha, centers, n_clusters, my_loss):
        super().__init__()
        self.alpha = alpha
        self.register_buffer('cluster_centers', torch.as_tensor(centers).float().cuda(),
                             persistent=False)
        self.n_clusters = n_clusters
        self.my_loss = my_loss

    def forward(self, ypred, ytrue):
        score, res = ypred[0], ypred[1]
        l1 = _kl(score, ytrue[0]) if self.use_kl else F.cross_entropy(score, ytrue[0])
        mode = _pose_mode_of(self.my_loss, res.shape[-1])
        if FUSE_EXPECTED_POSE and mode in (L.POSE_GEODESIC_AA, L.POSE_GEODESIC_Q) and \
                not getattr(self.my_loss, 'reduce', True) and score.is_cuda and not ytrue[1].requires_grad:
            # one fused launch instead of the reference's python loop over the K bins
            rows = ops.expected_pose_loss(score, res, ytrue[1], self.cluster_centers, mode)
            return l1 + self.alpha * rows.mean()
        if self.per_bin_delta:       # residual is [B, K, ndim]: one delta per bin (Multires)
            y = self.cluster_centers + res
            l2 = _expected_pose_loss(score, self.my_loss, ytrue[1], lambda k: y[:, k])
        else:
            l2 = _expected_pose_loss(score, self.my_loss, ytrue[1],
                                     lambda k: res + self.cluster_centers[k:k + 1])
        return l1 + self.alpha * l2


class RelaXedProbabilisticLoss(_ProbBase):
    """binDeltaLosses.py:111-129 (GMM means as centres)"""

    def __init__(self, alpha, gmm_file, my_loss):
        with open(gmm_file, 'rb') as f:
            gmm = pickle.load(f)
        super().__init__(alpha, gmm.means_, gmm.n_components, my_loss)


class ProbabilisticLoss(_ProbBase):
    """binDeltaLosses.py:132-150"""
    use_kl = False

    def __init__(self, alpha, kmeans_file, my_loss):
        km, centers = _load_centers(kmeans_file)
        super().__init__(alpha, centers, km.n_clusters, my_loss)


class RelaXedProbabilisticLossQ(_ProbBase):
    """binDeltaLosses.py:153-171"""

    def __init__(self, alpha, kmeans_file, my_loss):
        km, centers = _load_centers(kmeans_file)
        super().__init__(alpha, _quat.convert_dictionary(centers), km.n_clusters, my_loss)


class RelaXedProbabilisticMultiresLoss(RelaXedProbabilisticLoss):
    """binDeltaLosses.py:169-180 (the reference names the dictionary argument kmeans_file here)"""
    per_bin_delta = True

    def __init__(self, alpha, kmeans_file, my_loss):
        super().__init__(alpha, kmeans_file, my_loss)


class ProbabilisticMultiresLoss(ProbabilisticLoss):
    """binDeltaLosses.py:189-201"""
    per_bin_delta = True


class RelaXedProbabilisticMultiresLossQ(RelaXedProbabilisticLossQ):
    """binDeltaLosses.py:204-208"""
    per_bin_delta = True


class loss_m2(nn.Module):
    """CE + alpha * MSE(residual, ydata_res[b, argmax score_b, :]) — binDeltaLosses.py:280-297: the
    reference's `bmm(ytrue[1].permute(0, 2, 1), onehot[B, K, 1])` picks, per sample, the row of the
    per-bin residual targets ytrue[1] [B, K, ndim] that belongs to the PREDICTED bin."""

    def __init__(self, alpha, num_clusters):
        super().__init__()
        self.alpha = alpha
        self.num_clusters = num_clusters

    def forward(self, ypred, ytrue):
        lc, _, ind = ops.bd_loss(ypred[0], ytrue[0], None, None, None, L.POSE_NONE, False)
        yres = ytrue[1].gather(1, ind.view(-1, 1, 1).expand(-1, 1, ytrue[1].shape[2])).squeeze(1)
        lr = ops.pose_loss(ypred[1], yres, L.POSE_MSE)
        return lc + self.alpha * lr


class loss_m3(_ProbBase):
    """binDeltaLosses.py:300-320"""

    def __init__(self, alpha, kmeans_file, my_loss=None):
        km, centers = _load_centers(kmeans_file)
        super().__init__(alpha, centers, km.n_clusters, my_loss)


class loss_m4(loss_m3):
    """binDeltaLosses.py:323-334"""
    per_bin_delta = True
