# -*- coding: utf-8 -*-
"""Drop-in for the reference's helperFunctions module (helperFunctions.py:1-130): constants, image-name
parsing, Euler -> rotation matrix, the RBF gamma of the soft-bin generators, the cyclical-LR SGD of
the snapshot-ensemble evaluate* scripts and the per-class accuracy.

`get_gamma` runs on the device (bdp_min_key_gap); `mySGD.step()` updates every parameter tensor in ONE
launch (bdp_sgd_step) instead of one small elementwise kernel chain per tensor."""
import ctypes as C

import numpy as np
import torch
from torch.optim import Optimizer

from bdpose import _lib as L
from bdpose import ops

# helperFunctions.py:16-20
classes = ['aeroplane', 'bicycle', 'boat', 'bottle', 'bus', 'car', 'chair', 'diningtable', 'motorbike',
           'sofa', 'train', 'tvmonitor']
eps = 1e-6


def parse_name(image_name):
    """helperFunctions.py:24-33: '<synset>_<model>_a<az>_e<el>_t<ct>_d<dist>' ->
    (synset, model, az, el, ct, d); every numeric field carries a one-letter tag."""
    f = str(image_name).split('_', 5)
    if len(f) < 6:
        raise IndexError('list index out of range')       # what the reference raises on short names
    return (f[0], f[1]) + tuple(float(v[1:]) for v in f[2:6])


def rotation_matrix(az, el, ct):
    """helperFunctions.py:37-48: R = Rz(ct) . Rx(el) . Rz(az), angles in degrees.  One image at a
    time stays on the host; whole datasets go through bdpose.ops.euler_to_pose."""
    def rz(deg):
        c, s = np.cos(np.radians(deg)), np.sin(np.radians(deg))
        return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    cb, sb = np.cos(np.radians(el)), np.sin(np.radians(el))
    rx = np.array([[1, 0, 0], [0, cb, -sb], [0, sb, cb]])
    return np.dot(np.dot(rz(ct), rx), rz(az))


def get_gamma(kmeans_dict):
    """helperFunctions.py:51-58: 1 / (2 * min_i min_{j != i} ||k_i - k_j||^2)."""
    k = torch.as_tensor(np.ascontiguousarray(kmeans_dict), dtype=torch.float64).cuda()
    return 1.0 / (2.0 * float(ops.min_key_gap(k)))


class mySGD(Optimizer):
    """helperFunctions.py:62-120: SGD (optional momentum / weight decay / Nesterov) with the cyclical
    learning rate of snapshot ensembles: with t = ((step - 1) mod c + 1) / c the step size is
    (1 - 2t) alpha1 + 2t alpha2 for t <= 1/2 and 2(1 - t) alpha2 + (2t - 1) alpha1 above."""

    def __init__(self, params, c, alpha1=1e-6, alpha2=1e-8, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        defaults = dict(alpha1=alpha1, alpha2=alpha2, momentum=momentum, dampening=dampening,
                        weight_decay=weight_decay, nesterov=nesterov)
        super(mySGD, self).__init__(params, defaults)
        self.c = c

    def __setstate__(self, state):
        super(mySGD, self).__setstate__(state)
        for group in self.param_groups:
            group.setdefault('nesterov', False)

    def _step_size(self, step, group):
        t = (np.fmod(step - 1, self.c) + 1) / self.c
        if t <= 0.5:
            return (1 - 2 * t) * group['alpha1'] + 2 * t * group['alpha2']
        return 2 * (1 - t) * group['alpha2'] + (2 * t - 1) * group['alpha1']

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            momentum = group['momentum']
            rows, keep = [], []
            max_n = 0
            dev = None
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("mySGD: parameters must live on CUDA (this package has no CPU path)")
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = 0
                state['step'] += 1
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                    keep.append(g)
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("mySGD: parameters must be contiguous float32 tensors")
                first = 0
                buf_ptr = None
                if momentum != 0:
                    if 'momentum_buffer' not in state:
                        state['momentum_buffer'] = torch.empty_like(p)
                        first = 1
                    buf_ptr = state['momentum_buffer'].data_ptr()
                rows.append(L.SgdTensor(p.data_ptr(), g.data_ptr(), buf_ptr, p.numel(),
                                        float(self._step_size(state['step'], group)), first))
                max_n = max(max_n, p.numel())
                dev = p.device
            if not rows:
                continue
            table = (L.SgdTensor * len(rows))(*rows)
            raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(dev)
            with torch.cuda.device(dev):
                st = L.lib().bdp_sgd_step(C.c_void_p(raw.data_ptr()), len(rows), max_n,
                                          float(group['weight_decay']), float(momentum),
                                          float(group['dampening']), 1 if group['nesterov'] else 0,
                                          L.stream_ptr())
            L.check(st, "bdp_sgd_step")
            del keep
        return loss


def get_accuracy(ytrue, ypred, num_classes):
    """helperFunctions.py:123-130: mean over classes of the per-class recall."""
    acc = np.zeros(num_classes)
    for i in range(num_classes):
        acc[i] = np.sum((ytrue == i) * (ypred == i)) / np.sum(ytrue == i)
    return np.mean(acc)
