# -*- coding: utf-8 -*-
"""Drop-in for the reference's objectnetHelperFunctions module (objectnetHelperFunctions.py:1-231).

MODEL classes (110-231): ObjectNet3D heads take cat(features, onehot(label)) as input and use ONE
bin / res MLP pair for all 100 categories.  The pair runs as a two-head stack on the tcgen05 head
kernels.

DATASET classes (TrainImages / TestImages, 23-107): same constructor, item layout and
`shuffle_images()` as the reference; the pose targets, bins and residuals of the WHOLE dataset are
computed once at construction on the GPU (bdp_euler_to_pose + bdp_assign_nearest) instead of per item
in the DataLoader workers (sklearn predict on a [C,3] array per item).  Only the image read stays per
item (PIL, disk I/O — outside the kernel scope)."""
import os
import pickle

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


def _preprocess():
    """objectnetHelperFunctions.py:19-20 (built lazily: torchvision is only needed for image items)"""
    from torchvision import transforms
    normalize = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    return transforms.Compose([transforms.Resize([224, 224]), transforms.ToTensor(), normalize])


def __getattr__(name):
    if name == 'preprocess':        # module-level transform of the reference (PEP 562, lazy)
        return _preprocess()
    raise AttributeError("module 'objectnetHelperFunctions' has no attribute %r" % name)


def _load_image_lists(data_path, classes):
    import scipy.io as spio
    out = []
    for c in classes:
        tmp = spio.loadmat(os.path.join(data_path, c + '_info'), squeeze_me=True)
        out.append(np.atleast_1d(tmp['image_names']))
    return out


class TrainImages(torch.utils.data.Dataset):
    """objectnetHelperFunctions.py:23-66: item = one image per class; ydata [C,3] fp32, label [C,1],
    ydata_bin [C] int64 = kmeans.predict(ydata), ydata_res [C,3] = ydata - centers[bin]."""

    def __init__(self, data_path, classes, dict_size=16):
        import binDeltaGenerators as G
        self.db_path = data_path
        self.classes = classes
        self.num_classes = len(self.classes)
        self.list_image_names = _load_image_lists(data_path, classes)
        self.num_images = np.array([len(self.list_image_names[i]) for i in range(self.num_classes)])
        self.image_names = self.list_image_names
        kmeans_file = 'data/kmeans_dictionary_axis_angle_' + str(dict_size) + '.pkl'
        with open(kmeans_file, 'rb') as f:
            self.kmeans = pickle.load(f)
        # every label of the dataset in two launches (float32 targets, as the per-item code: 54-60)
        per_class = G.pose_targets_from_names(self.list_image_names, 1.0, False, torch.float32)
        y = torch.from_numpy(np.concatenate(per_class))
        b, r = G.assign_labels(y, np.asarray(self.kmeans.cluster_centers_))
        off = np.concatenate([[0], np.cumsum(self.num_images)])
        b, r = b.cpu(), r.cpu()
        self._y = [y[off[i]:off[i + 1]] for i in range(self.num_classes)]
        self._bin = [b[off[i]:off[i + 1]] for i in range(self.num_classes)]
        self._res = [r[off[i]:off[i + 1]] for i in range(self.num_classes)]
        self._index = [{n: j for j, n in enumerate(names)} for names in self.list_image_names]
        self._pre = None

    def __len__(self):
        return np.amax(self.num_images)

    def _image(self, cls, image_name):
        from PIL import Image
        if self._pre is None:
            self._pre = _preprocess()
        return self._pre(Image.open(os.path.join(self.db_path, self.classes[cls], image_name + '.png')))

    def __getitem__(self, idx):
        xdata, rows, label = [], [], []
        for i in range(self.num_classes):
            image_name = self.image_names[i][idx % self.num_images[i]]
            label.append(i * torch.ones(1).long())
            xdata.append(self._image(i, image_name))
            rows.append(self._index[i][image_name])
        return {'xdata': torch.stack(xdata),
                'ydata': torch.stack([self._y[i][j] for i, j in enumerate(rows)]),
                'label': torch.stack(label),
                'ydata_bin': torch.stack([self._bin[i][j] for i, j in enumerate(rows)]),
                'ydata_res': torch.stack([self._res[i][j] for i, j in enumerate(rows)])}

    def shuffle_images(self):
        self.image_names = [np.random.permutation(self.list_image_names[i]) for i in range(self.num_classes)]


class TestImages(torch.utils.data.Dataset):
    """objectnetHelperFunctions.py:69-107: item = one image; the bin and the residual are computed
    from the float64 pose target (100-101), ydata is its float32 copy."""

    def __init__(self, data_path, classes, dict_size=16):
        import binDeltaGenerators as G
        self.db_path = data_path
        self.classes = classes
        self.num_classes = len(self.classes)
        self.list_image_names = _load_image_lists(data_path, classes)
        self.list_labels = [i * np.ones(len(n), dtype='int') for i, n in enumerate(self.list_image_names)]
        self.image_names = np.concatenate(self.list_image_names)
        self.labels = np.concatenate(self.list_labels)
        kmeans_file = 'data/kmeans_dictionary_axis_angle_' + str(dict_size) + '.pkl'
        with open(kmeans_file, 'rb') as f:
            self.kmeans = pickle.load(f)
        y64 = torch.from_numpy(G.pose_targets_from_names([self.image_names], 1.0, False, torch.float64)[0])
        b, r = G.assign_labels(y64, np.asarray(self.kmeans.cluster_centers_))
        self._y, self._bin, self._res = y64.float(), b.cpu(), r.cpu()
        self._pre = None

    def __len__(self):
        return len(self.image_names)

    def __getitem__(self, idx):
        from PIL import Image
        if self._pre is None:
            self._pre = _preprocess()
        image_name = self.image_names[idx]
        label = self.labels[idx]
        img = Image.open(os.path.join(self.db_path, self.classes[label], image_name + '.png'))
        return {'xdata': self._pre(img), 'ydata': self._y[idx],
                'label': int(label) * torch.ones(1).long(),
                'ydata_bin': self._bin[idx].reshape(1), 'ydata_res': self._res[idx]}


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
