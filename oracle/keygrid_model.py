"""numpy model of the key-grid candidate filter — TEST INFRASTRUCTURE ONLY (the product never imports
this module).

It restates, in float64, the two tests the CUDA build applies to a (cell box, key) pair
(`csrc/assign.cu`: `box_bounds` / `bisector_min` / `keygrid_cell_kernel`) so that the claim the pruned
query rests on can be checked on the CPU, independently of the GPU parity tests:

    every key that is nearest to SOME point of a cell is in the cell's candidate list

A key k stays in the list of box B when
  (1) mindist^2(B, c_k) <= min_j maxdist^2(B, c_j)          (k is not farther than the pivot everywhere)
  (2) min over B of |x - c_k|^2 - |x - c_p|^2 <= 0          (k beats the pivot key p at some corner; the
                                                            difference is linear in x)
— both are necessary conditions for "k is nearest somewhere in B", so the list is a superset of the
keys the brute-force scan (binDeltaGenerators.py:27, sklearn `predict`) can return inside B.
"""
import numpy as np


def cell_boxes(lo, hi, G):
    """Boxes [n_cells, d] (lower, upper corner) of the uniform G^d grid over the box [lo, hi]; cell index
    = sum_k coordinate_k * G^k (first axis fastest), as the query maps points to cells."""
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    d = lo.size
    cell = (hi - lo) / G
    n = np.arange(G ** d)
    idx = np.stack([(n // G ** k) % G for k in range(d)], axis=1)
    blo = lo + idx * cell
    return blo, blo + cell


def candidate_mask(centers, blo, bhi, slack=0.0):
    """[n_boxes, K] bool: keys the build keeps for every box (tests (1) and (2) with thresholds relaxed
    by `slack` in squared-distance units — the CUDA build relaxes them by its fp32 rounding)."""
    c = np.asarray(centers, dtype=np.float64)
    a = c[None, :, :] - blo[:, None, :]
    b = bhi[:, None, :] - c[None, :, :]
    far = np.maximum(np.abs(a), np.abs(b))
    near = np.maximum(np.maximum(-a, -b), 0.0)
    maxd2 = (far ** 2).sum(2)
    mind2 = (near ** 2).sum(2)
    piv = maxd2.argmin(1)
    u = maxd2[np.arange(blo.shape[0]), piv]
    cp = c[piv]
    dlt = c[None, :, :] - cp[:, None, :]
    corner = np.where(dlt > 0, bhi[:, None, :], blo[:, None, :])
    f = (-2.0 * dlt * corner).sum(2) + (c ** 2).sum(1)[None, :] - (cp ** 2).sum(1)[:, None]
    return (mind2 <= u[:, None] + slack) & (f <= slack)


def point_cells(x, lo, hi, G):
    """Cell index of every point (floor of the scaled coordinate, clipped to the grid)."""
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    d = lo.size
    t = np.floor((np.asarray(x, dtype=np.float64) - lo) / ((hi - lo) / G)).astype(np.int64)
    t = np.clip(t, 0, G - 1)
    return (t * (G ** np.arange(d))).sum(1)
