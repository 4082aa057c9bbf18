"""Evaluation reductions on the device (axisAngle.get_error/get_error2, quaternion.get_error/
get_error2 — reference: axisAngle.py:45-95, quaternion.py:33-76)."""
import numpy as np
import torch

from . import ops


def _to_dev(a, device=None):
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.as_tensor(np.ascontiguousarray(a))
    if not t.is_cuda:
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    return t


def get_error(ygt, yhat, quaternion=False, verbose=True):
    """(acc, medErr, errors[N] numpy) exactly as the reference returns them."""
    ygt, yhat = _to_dev(ygt), _to_dev(yhat)
    err = ops.geodesic_error_deg(ygt, yhat, quaternion=quaternion)
    med, cnt, b30, mx = ops.error_stats(err, None, 1)
    n = err.numel()
    host = torch.cat([med, mx, b30.double()]).tolist()
    medErr, maxErr = host[0], host[1]
    acc = 100 * host[2] / n if n > 0 else float("nan")
    if verbose:
        print('Error stats- Median: {0}, Max: {1}, <30: {2}'.format(medErr, maxErr, acc))
    return acc, medErr, err.cpu().numpy()


def get_error2(ygt, yhat, labels, num, quaternion=False):
    """mean over classes of the per-class median error (NaN if a class is empty, like np.median)."""
    ygt, yhat = _to_dev(ygt), _to_dev(yhat)
    labels = _to_dev(labels).reshape(-1)
    err = ops.geodesic_error_deg(ygt, yhat, quaternion=quaternion)
    med, _, _, _ = ops.error_stats(err, labels, int(num))
    return float(med.mean())


# ---- detection-metric helpers (ports of the reference's MATLAB functions) ----------------------------
def box_overlap(a, b):
    """box_overlap.m: symmetric IoU between every box of a [n,4] and the single box b [4], boxes as
    [x1 y1 x2 y2] with inclusive pixel coordinates (+1 on widths / heights); 0 where they are disjoint."""
    a = np.atleast_2d(np.asarray(a, dtype=np.float64))
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    w = np.minimum(a[:, 2], b[2]) - np.maximum(a[:, 0], b[0]) + 1
    h = np.minimum(a[:, 3], b[3]) - np.maximum(a[:, 1], b[1]) + 1
    inter = w * h
    aarea = (a[:, 2] - a[:, 0] + 1) * (a[:, 3] - a[:, 1] + 1)
    barea = (b[2] - b[0] + 1) * (b[3] - b[1] + 1)
    o = inter / (aarea + barea - inter)
    o[(w <= 0) | (h <= 0)] = 0
    return o


def VOCap(rec, prec):
    """VOCap.m: area under the monotone envelope of the precision/recall curve (VOC 2010+)."""
    mrec = np.concatenate([[0.0], np.asarray(rec, dtype=np.float64).reshape(-1), [1.0]])
    mpre = np.concatenate([[0.0], np.asarray(prec, dtype=np.float64).reshape(-1), [0.0]])
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    i = np.nonzero(mrec[1:] != mrec[:-1])[0] + 1
    return float(np.sum((mrec[i] - mrec[i - 1]) * mpre[i]))
