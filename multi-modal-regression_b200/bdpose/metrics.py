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
