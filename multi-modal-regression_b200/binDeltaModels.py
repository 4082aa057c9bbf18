# -*- coding: utf-8 -*-
"""Drop-in for the reference's binDeltaModels module (binDeltaModels.py:1-178) on the B200 head
kernels.

* `bin_3layer` / `res_3layer` keep the reference's layer names (fc1, bn1, fc2, bn2, fc3), so
  state_dict keys, optimizers and `model.bin_models[i](x)` calls are unchanged; a single module runs
  as a one-head stack through the same tcgen05 GEMM / BatchNorm / fc3 kernels.
* `OneBinDeltaModel.forward(x, label)` runs all 2*C heads of the model fused: one flattened fc1 GEMM,
  one grouped fc2 GEMM, BatchNorm on feature-major activations, label-selected fc3 — with the one-hot
  built on the device (the reference round-trips it through the CPU every forward,
  binDeltaModels.py:116-117).  `forward_mixed(x, mix)` is the soft-mixing form the joint
  category+pose scripts build by hand (learnJointCatPoseModel_weighted.py:107-115).
* `OneDeltaPerBinModel` / `ProbabilisticOneDeltaPerBinModel` run their C*K two-layer
  delta heads as one stack too (`bdpose.head.Mlp2Stack`: one fc1 GEMM over 2048 -> C*K*N3, BatchNorm,
  a batched output layer) instead of the reference's python loop over C*K modules; a single
  1- / 2-layer block called on its own stays on stock torch layers.
"""
import torch
from torch import nn
import torch.nn.functional as F

from bdpose import head as _head


def _feature_model(feature_network):
    import featureModels
    if feature_network == 'resnet':
        return featureModels.resnet_model('resnet50', 'layer4').cuda()
    if feature_network == 'vgg':
        return featureModels.vgg_model('vgg13', 'fc6').cuda()
    return None    # the reference silently leaves feature_model undefined for other strings


# ---- building blocks ----------------------------------------------------------------------------------
class bin_1layer(nn.Module):
    """binDeltaModels.py:16-23"""

    def __init__(self, N0, num_clusters):
        super().__init__()
        self.fc = nn.Linear(N0, num_clusters)

    def forward(self, x):
        return self.fc(x)


class res_1layer(nn.Module):
    """binDeltaModels.py:26-33"""

    def __init__(self, N0, ndim):
        super().__init__()
        self.fc = nn.Linear(N0, ndim)

    def forward(self, x):
        return self.fc(x)


class _mlp2(nn.Module):
    """fc2(relu(bn1(fc1 x))) — bin_2layer / res_2layer, binDeltaModels.py:36-59"""

    def __init__(self, N0, N1, Nout):
        super().__init__()
        self.fc1 = nn.Linear(N0, N1, bias=False)
        self.bn1 = nn.BatchNorm1d(N1)
        self.fc2 = nn.Linear(N1, Nout)

    def forward(self, x):
        # the C*K sibling delta heads of OneDeltaPerBinModel, called one by one by a script-defined
        # forward (learnJointCatPoseModel_weighted.py:117-118), run as one fused two-layer stack
        fam = self.__dict__.get('_family')
        if fam is not None and _head.MEMO and x.is_cuda:
            return fam.output_of(self, x, self.training)
        return self.fc2(F.relu(self.bn1(self.fc1(x))))

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new


class bin_2layer(_mlp2):
    def __init__(self, N0, N1, num_clusters):
        super().__init__(N0, N1, num_clusters)


class res_2layer(_mlp2):
    def __init__(self, N0, N1, ndim):
        super().__init__(N0, N1, ndim)


class _mlp3(nn.Module):
    """fc3(relu(bn2(fc2(relu(bn1(fc1 x)))))) with fc1/fc2 bias-free — binDeltaModels.py:62-91."""

    def __init__(self, N0, N1, N2, Nout):
        super().__init__()
        self.fc1 = nn.Linear(N0, N1, bias=False)
        self.bn1 = nn.BatchNorm1d(N1)
        self.fc2 = nn.Linear(N1, N2, bias=False)
        self.bn2 = nn.BatchNorm1d(N2)
        self.fc3 = nn.Linear(N2, Nout)
        object.__setattr__(self, '_solo', None)

    def forward(self, x):
        # scripts call `self.bin_models[i](x)` head by head (learnJointCatPoseModel_weighted.py:112-113):
        # the family runs all sibling heads fused on the first call and hands out slices
        fam = self.__dict__.get('_family')
        if fam is not None and _head.MEMO and x.is_cuda:
            return fam.output_of(self, x, self.training)
        # a head on its own = a stack of one, mixing weight 1
        solo = self._solo
        if solo is None or solo.heads[0] is not self:
            solo = _head.HeadStack([[self]])
            object.__setattr__(self, '_solo', solo)
        ones = torch.ones(x.shape[0], 1, device=x.device)
        return _head.run_heads(solo, x, ones, self.training)[0]

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == '_solo':
                continue
            if k == '_family':           # re-created by the copied model (or copied with its lists)
                new.__dict__[k] = copy.deepcopy(v, memo)
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        object.__setattr__(new, '_solo', None)
        return new


class bin_3layer(_mlp3):
    def __init__(self, N0, N1, N2, num_clusters):
        super().__init__(N0, N1, N2, num_clusters)


class res_3layer(_mlp3):
    def __init__(self, N0, N1, N2, ndim):
        super().__init__(N0, N1, N2, ndim)


# ---- models ---------------------------------------------------------------------------------------------
class OneBinDeltaModel(nn.Module):
    """binDeltaModels.py:99-121.  forward(x, label) -> [y1 [B, num_clusters], y2 [B, ndim]]."""

    def __init__(self, feature_network, num_classes, num_clusters, N0, N1, N2, ndim):
        super().__init__()
        self.num_classes = num_classes
        self.num_clusters = num_clusters
        self.ndim = ndim
        fm = _feature_model(feature_network)
        if fm is not None:
            self.feature_model = fm
        self.bin_models = nn.ModuleList([bin_3layer(N0, N1, N2, num_clusters) for i in range(self.num_classes)]).cuda()
        self.res_models = nn.ModuleList([res_3layer(N0, N1, N2, ndim) for i in range(self.num_classes)]).cuda()
        # the heads know their siblings: a script-defined forward that calls them one by one
        # (learnJointCatPoseModel_weighted.py:107-126) still runs them as one fused stack
        fam = _head.HeadFamily([self.bin_models, self.res_models])
        for m in list(self.bin_models) + list(self.res_models):
            object.__setattr__(m, '_family', fam)

    def _heads(self):
        bins, ress = list(self.bin_models), list(self.res_models)
        fam = bins[0].__dict__.get('_family') if bins else None
        if fam is None or len(fam.lists) != 2 or fam.lists[0] is not self.bin_models or \
                fam.lists[1] is not self.res_models:
            fam = _head.HeadFamily([self.bin_models, self.res_models])
            for m in bins + ress:
                object.__setattr__(m, '_family', fam)
        return fam.stack()

    def stacked_head_parameters(self):
        """Opt-in fast path for new training loops: the 2*C heads' weights as 10 stacked
        nn.Parameters (shared memory with the per-module ones).  Build the optimizer over
        `list(model.feature_model.parameters()) + model.stacked_head_parameters()`; backward() then
        leaves the head gradients on the stacked Parameters only."""
        return self._heads().stacked_parameters()

    def forward_features(self, feat, label=None, mix=None):
        """Heads only, on precomputed features [B, N0]: label [B,1] int64 or mix [B, C] weights."""
        if mix is None:
            mix = _head.onehot(label, self.num_classes)
        y1, y2 = _head.run_heads(self._heads(), feat, mix, self.training)
        return [y1, y2]

    def forward_mixed(self, x, mix):
        """Soft category mixing (learnJointCatPoseModel_weighted.py:107-115) with gradient to mix."""
        return self.forward_features(self.feature_model(x), mix=mix)

    def forward(self, x, label):
        x = self.feature_model(x)
        return self.forward_features(x, label=label)


def _onehot_cpu_like(idx, n):
    return torch.zeros(idx.size(0), n, device=idx.device).scatter_(1, idx, 1.0)


class OneDeltaPerBinModel(nn.Module):
    """binDeltaModels.py:124-151: C bin heads (fused 3-layer stack) + C*K res_2layer heads (fused
    2-layer stack), delta selected by class then by argmax bin."""

    def __init__(self, feature_network, num_classes, num_clusters, N0, N1, N2, N3, ndim):
        super().__init__()
        self.num_classes = num_classes
        self.num_clusters = num_clusters
        self.ndim = ndim
        fm = _feature_model(feature_network)
        if fm is not None:
            self.feature_model = fm
        self.bin_models = nn.ModuleList([bin_3layer(N0, N1, N2, num_clusters) for i in range(self.num_classes)]).cuda()
        self.res_models = nn.ModuleList([res_2layer(N0, N3, ndim) for i in range(self.num_classes * self.num_clusters)]).cuda()
        # sibling families: a script-defined forward that calls the heads one by one
        # (learnJointCatPoseModel_weighted.py:116-118) still runs each list as one fused stack
        self._families()

    def _families(self):
        bins, ress = list(self.bin_models), list(self.res_models)
        fb = bins[0].__dict__.get('_family') if bins else None
        if fb is None or len(fb.lists) != 1 or fb.lists[0] is not self.bin_models:
            fb = _head.HeadFamily([self.bin_models])
            for m in bins:
                object.__setattr__(m, '_family', fb)
        fr = ress[0].__dict__.get('_family') if ress else None
        if fr is None or len(fr.lists) != 1 or fr.lists[0] is not self.res_models:
            fr = _head.HeadFamily([self.res_models], two_layer=True)
            for m in ress:
                object.__setattr__(m, '_family', fr)
        return fb, fr

    def _bin_scores(self, x, mix):
        return _head.run_heads(self._families()[0].stack(), x, mix, self.training)[0]

    def _class_deltas(self, x, class_label):
        """All K per-bin deltas of every sample's own class: [B, K, ndim].  The C*K two-layer heads
        run fused (one fc1 GEMM over the stacked 2048 -> C*K*N3 weights, BatchNorm, batched output
        layer); the reference's one-hot matmul select (binDeltaModels.py:141-145) is a gather."""
        st = self._families()[1].stack()
        y = _head.run_mlp2_all(st, x, self.training)                     # [B, C*K, ndim]
        y = y.view(y.shape[0], self.num_classes, self.num_clusters, self.ndim)
        idx = class_label.reshape(-1, 1, 1, 1).expand(-1, 1, self.num_clusters, self.ndim)
        return torch.gather(y, 1, idx).squeeze(1)

    def forward(self, x, class_label):
        x = self.feature_model(x)
        class_mix = _onehot_cpu_like(class_label, self.num_classes)
        y1 = self._bin_scores(x, class_mix)
        y2 = self._class_deltas(x, class_label)                          # [B, K, ndim]
        pose_label = torch.argmax(y1, dim=1, keepdim=True)               # first maximum, as torch.max
        y2 = torch.gather(y2, 1, pose_label.unsqueeze(2).expand(-1, 1, self.ndim)).squeeze(1)
        return [y1, y2]


class ProbabilisticOneDeltaPerBinModel(OneDeltaPerBinModel):
    """binDeltaModels.py:154-178: as above but returns all K deltas, [B, K, ndim]."""

    def forward(self, x, class_label):
        x = self.feature_model(x)
        class_mix = _onehot_cpu_like(class_label, self.num_classes)
        y1 = self._bin_scores(x, class_mix)
        return [y1, self._class_deltas(x, class_label)]
