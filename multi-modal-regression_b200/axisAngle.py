# -*- coding: utf-8 -*-
"""Drop-in for the reference's axisAngle module (axisAngle.py:1-120) on the B200 kernels.

get_y / get_R act on ONE 3x3 matrix / 3-vector on the host exactly as the reference's helpers do
(they are called per file name while a dataset is being indexed); everything that is batched —
the error metrics and the geodesic loss — runs in the CUDA library.
"""
import numpy as np
from torch import nn

from helperFunctions import eps
from bdpose import ops, metrics
from bdpose import _lib as L


def get_y(R):
    """Axis-angle log map of one rotation matrix (reference axisAngle.py:19-29): the zero vector
    when the extracted axis has norm <= eps (includes theta = pi)."""
    R = np.asarray(R)
    t = np.arccos(np.clip(0.5 * (np.trace(R) - 1), -1., 1.))
    S = 0.5 * (R - R.T)
    v = np.array([S[2, 1], S[0, 2], S[1, 0]])
    n = np.linalg.norm(v)
    v = v / n if n > eps else np.zeros(3)
    return t * v


def get_R(v):
    """Rodrigues exp map of one axis-angle vector (reference axisAngle.py:33-41)."""
    t = np.linalg.norm(v)
    if t < eps:
        return np.eye(3)
    a = v / t
    V = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(t) * V + (1 - np.cos(t)) * np.dot(V, V)


def get_error(ygt, yhat):
    """(acc@30deg in %, median error, per-sample errors in degrees) — axisAngle.py:45-66."""
    return metrics.get_error(ygt, yhat, quaternion=False)


def get_error2(ygt, yhat, labels, num):
    """Mean over `num` classes of the per-class median error — axisAngle.py:70-95."""
    return metrics.get_error2(ygt, yhat, labels, num, quaternion=False)


class geodesic_loss(nn.Module):
    """theta = 2 acos(clamp(|cos(at/2)cos(ap/2) + sin(at/2)sin(ap/2) <t^,p^>|)) — axisAngle.py:103-120,
    forward and hand-derived backward in one fused launch."""

    def __init__(self, reduce=True):
        super().__init__()
        self.eps = eps
        self.reduce = reduce

    def forward(self, ypred, ytrue):
        return ops.pose_loss(ypred, ytrue, L.POSE_GEODESIC_AA, reduce=self.reduce)
