# -*- coding: utf-8 -*-
"""Drop-in for the MODEL classes of the reference's objectnetHelperFunctions module
(objectnetHelperFunctions.py:110-231): ObjectNet3D heads take cat(features, onehot(label)) as input
and use ONE bin / res MLP pair for all 100 categories.  The pair runs as a two-head stack on the
tcgen05 head kernels.  The dataset classes of that module (TrainImages / TestImages: disk I/O) stay
the reference's own."""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from bdpose import head as _head
import binDeltaModels as _bdm
from binDeltaModels import _feature_model


# the reference's ObjectNet copies of the building blocks use lower-case argument names
class bin_3layer(_bdm.bin_3layer):
    """objectnetHelperFunctions.py:110-123"""

    def __init__(self, n0, n1, n2, num_clusters):
        super().__init__(n0, n1, n2, num_clusters)


class res_3layer(_bdm.res_3layer):
    """objectnetHelperFunctions.py:126-139"""

    def __init__(self, n0, n1, n2, dim):
        super().__init__(n0, n1, n2, dim)


class res_2layer(_bdm.res_2layer):
    """objectnetHelperFunctions.py:142-152"""

    def __init__(self, n0, n1, dim):
        super().__init__(n0, n1, dim)


def _cat_onehot(feat, label, num_classes):
    return torch.cat((feat, _head.onehot(label, num_classes)), dim=1)


class OneBinDeltaModel(nn.Module):
    """objectnetHelperFunctions.py:155-172"""

    def __init__(self, num_classes, dict_size=200, n0=2048, n1=1000, n2=500, dim=3):
        super().__init__()
        self.num_classes = num_classes
        self.num_clusters = dict_size
        fm = _feature_model('resnet')
        if fm is not None:
            self.feature_model = fm
        self.bin_model = bin_3layer(n0 + num_classes, n1, n2, self.num_clusters).cuda()
        self.res_model = res_3layer(n0 + num_classes, n1, n2, dim).cuda()
        object.__setattr__(self, '_stack', None)

    def forward_features(self, feat, label):
        st = self.__dict__.get('_stack')
        if st is None or st.heads[0] is not self.bin_model or st.heads[1] is not self.res_model:
            st = _head.HeadStack([[self.bin_model], [self.res_model]])
            object.__setattr__(self, '_stack', st)
        x = _cat_onehot(feat, label, self.num_classes)
        ones = torch.ones(x.shape[0], 1, device=x.device)
        y1, y2 = _head.run_heads(st, x, ones, self.training)
        return [y1, y2]

    def forward(self, x, label):
        return self.forward_features(self.feature_model(x), label)


class RegressionModel(nn.Module):
    """objectnetHelperFunctions.py:201-215: pi * tanh(res_3layer(cat(features, onehot)))"""

    def __init__(self, num_classes, n0=2048, n1=1000, n2=500, dim=3):
        super().__init__()
        self.num_classes = num_classes
        fm = _feature_model('resnet')
        if fm is not None:
            self.feature_model = fm
        self.pose_model = res_3layer(n0 + num_classes, n1, n2, dim).cuda()

    def forward(self, x, label):
        x = _cat_onehot(self.feature_model(x), label, self.num_classes)
        return np.pi * torch.tanh(self.pose_model(x))


class ClassificationModel(nn.Module):
    """objectnetHelperFunctions.py:218-231"""

    def __init__(self, num_classes, dict_size=16, n0=2048, n1=1000, n2=500):
        super().__init__()
        self.num_classes = num_classes
        fm = _feature_model('resnet')
        if fm is not None:
            self.feature_model = fm
        self.pose_model = bin_3layer(n0 + num_classes, n1, n2, dict_size).cuda()

    def forward(self, x, label):
        return self.pose_model(_cat_onehot(self.feature_model(x), label, self.num_classes))


class OneDeltaPerBinModel(nn.Module):
    """objectnetHelperFunctions.py:175-198: one bin head + dict_size res_2layer heads (fused two-layer
    stack), delta picked by the argmax bin."""

    def __init__(self, num_classes, dict_size=16, n0=2048, n1=1000, n2=500, n3=100, dim=3):
        super().__init__()
        self.ndim = dim
        self.num_classes = num_classes
        self.num_clusters = dict_size
        fm = _feature_model('resnet')
        if fm is not None:
            self.feature_model = fm
        self.bin_model = bin_3layer(n0 + num_classes, n1, n2, self.num_clusters).cuda()
        self.res_models = nn.ModuleList([res_2layer(n0 + num_classes, n3, dim) for i in range(self.num_clusters)]).cuda()

    def forward(self, x, label):
        x = _cat_onehot(self.feature_model(x), label, self.num_classes)
        y1 = self.bin_model(x)
        # the K per-bin delta heads run as one fused two-layer stack (the reference loops over them,
        # objectnetHelperFunctions.py:190); the one-hot bmm select (192-195) is a gather
        st = self.__dict__.get('_stack2')
        heads = list(self.res_models)
        if st is None or len(st.heads) != len(heads) or any(a is not b for a, b in zip(st.heads, heads)):
            st = _head.Mlp2Stack(heads)
            object.__setattr__(self, '_stack2', st)
        y2 = _head.run_mlp2_all(st, x, self.training)                             # [B, K, dim]
        pose = torch.argmax(y1, dim=1, keepdim=True)
        return [y1, torch.gather(y2, 1, pose.unsqueeze(2).expand(-1, 1, self.ndim)).squeeze(1)]
