# -*- coding: utf-8 -*-
"""Drop-in for the reference's quaternion module (quaternion.py:1-163) on the B200 kernels.
Quaternion layout is (w, x, y, z) with w = cos(theta/2)."""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from helperFunctions import eps
from bdpose import ops, metrics
from bdpose import _lib as L


def get_y(R):
    """Unit quaternion of one rotation matrix (reference quaternion.py:18-29); identity quaternion
    when the extracted axis has norm <= eps."""
    R = np.asarray(R)
    t = np.arccos(np.clip(0.5 * (np.trace(R) - 1), -1., 1.))
    S = 0.5 * (R - R.T)
    v = np.array([S[2, 1], S[0, 2], S[1, 0]])
    n = np.linalg.norm(v)
    if n > eps:
        v = v / n
    else:
        t, v = 0, np.zeros(3)
    h = t / 2.
    return np.array([np.cos(h), np.sin(h) * v[0], np.sin(h) * v[1], np.sin(h) * v[2]])


def get_error(ygt, yhat):
    """(acc@30deg, median, errors) with error = 2 acos|<q1,q2>| in degrees — quaternion.py:33-51."""
    return metrics.get_error(ygt, yhat, quaternion=True)


def get_error2(ygt, yhat, labels, num):
    """Mean of per-class medians — quaternion.py:55-76."""
    return metrics.get_error2(ygt, yhat, labels, num, quaternion=True)


def convert_dictionary(axisangle_dict):
    """Axis-angle dictionary [K,3] -> unit-quaternion dictionary [K,4] float64 numpy
    (quaternion.py:79-92), computed on the device."""
    aa = torch.as_tensor(np.ascontiguousarray(axisangle_dict), dtype=torch.float64).cuda()
    _, quat = ops.convert_axis_angle(aa, want_rot=False, want_quat=True)
    return quat.cpu().numpy()


class model_3layer(nn.Module):
    """2048->N1->N2->4 regressor with tanh + L2 normalisation (quaternion.py:101-115)."""

    def __init__(self, N0, N1, N2):
        super().__init__()
        self.fc1 = nn.Linear(N0, N1, bias=False)
        self.bn1 = nn.BatchNorm1d(N1)
        self.fc2 = nn.Linear(N1, N2, bias=False)
        self.bn2 = nn.BatchNorm1d(N2)
        self.fc3 = nn.Linear(N2, 4)

    def forward(self, x):
        x = F.relu(self.bn1(self.fc1(x)))
        x = F.relu(self.bn2(self.fc2(x)))
        return F.normalize(torch.tanh(self.fc3(x)))


class model_2layer(nn.Module):
    """quaternion.py:118-130"""

    def __init__(self, N0, N1):
        super().__init__()
        self.fc1 = nn.Linear(N0, N1, bias=False)
        self.bn1 = nn.BatchNorm1d(N1)
        self.fc2 = nn.Linear(N1, 4)

    def forward(self, x):
        x = F.relu(self.bn1(self.fc1(x)))
        return F.normalize(torch.tanh(self.fc2(x)))


class model_1layer(nn.Module):
    """quaternion.py:133-141"""

    def __init__(self, N0):
        super().__init__()
        self.fc = nn.Linear(N0, 4)

    def forward(self, x):
        return F.normalize(torch.tanh(self.fc(x)))


class geodesic_loss(nn.Module):
    """theta = 2 acos(clamp(|<ytrue, normalize(ypred)>|, +-(1-eps))) — quaternion.py:149-163."""

    def __init__(self, reduce=True):
        super().__init__()
        self.eps = eps
        self.reduce = reduce

    def forward(self, ypred, ytrue):
        return ops.pose_loss(ypred, ytrue, L.POSE_GEODESIC_Q, reduce=self.reduce)
