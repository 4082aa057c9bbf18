"""CPU stand-ins for bdp_kmeans_lloyd_step / bdp_kmeans_finalize — TEST INFRASTRUCTURE ONLY.

They follow the C-ABI contracts of include/bdpose.h (two-limb int64 fixed-point accumulators
[K, 2d+1], {changed} counter, inertia; finalisation rules of scikit-learn's _average_centers) in
numpy so that the HOST loop of bdpose.kmeans.kmeans_lloyd — sharding, the per-iteration all-reduce,
convergence tests, empty-cluster relocation — can be driven on CPU tensors under the gloo backend
(tests/test_dist_gloo.py).  The product never imports this module.
"""
import numpy as np
import torch

import bdpose_oracle as O


def lloyd_step(x, centers, state, fix_hi_bits, update=True, grid=None, want_inertia=False):
    X, C = x.numpy(), centers.numpy()
    K, d = C.shape
    lab = O.e_step(X, C).astype(np.int32)             # sklearn's E-step (oracle)
    old = state.labels.numpy()
    state.stats[0] += int((old != lab).sum())
    state.labels.copy_(torch.from_numpy(lab))
    state.inertia += float(((X - C[lab]) ** 2).sum())
    if update:
        xs = X * float(2.0 ** fix_hi_bits)
        f = np.floor(xs)
        hi = f.astype(np.int64)
        lo = np.trunc((xs - f) * 4294967296.0).astype(np.int64)
        acc = state.acc.view(K, 2 * d + 1).numpy()     # shares memory with the tensor
        for k in range(d):
            np.add.at(acc[:, 2 * k], lab, hi[:, k])
            np.add.at(acc[:, 2 * k + 1], lab, lo[:, k])
        np.add.at(acc[:, 2 * d], lab, 1)


def finalize(state, centers_old, centers_new, fix_hi_bits):
    K, d = centers_old.shape
    acc = state.acc.view(K, 2 * d + 1).numpy()
    cnt = acc[:, 2 * d]
    sums = np.empty((K, d))
    for k in range(d):
        for j in range(K):      # exact integer arithmetic, one rounding
            v = (int(acc[j, 2 * k]) << 32) + int(acc[j, 2 * k + 1])
            sums[j, k] = float(v) * 2.0 ** -(fix_hi_bits + 32) if abs(v) < 2 ** 1000 else np.inf
    big = int(np.argmax(cnt))
    new = np.empty((K, d))
    for j in range(K):
        if cnt[j] > 0:
            new[j] = sums[j] * (1.0 / cnt[j])
        elif big < j and cnt[big] > 0:
            new[j] = sums[big] * (1.0 / cnt[big])
        else:
            new[j] = sums[big]
    centers_new.copy_(torch.from_numpy(new))
    state.shift2[0] = float(((new - centers_old.numpy()) ** 2).sum())
    state.n_empty[0] = int((cnt == 0).sum())


BACKEND = (lloyd_step, finalize)
