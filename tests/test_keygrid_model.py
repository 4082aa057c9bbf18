"""CPU property test of the claim the pruned nearest-key query rests on (oracle/keygrid_model.py, a
float64 restatement of the build's two box tests): the candidate list of a cell contains every key the
brute-force scan (the oracle's `predict`, binDeltaGenerators.py:27) returns for a point of that cell —
for uniform rotations, for a clustered dictionary, with duplicate keys (ties) and in 4-D."""
import numpy as np
import pytest

import bdpose_oracle as O
import keygrid_model as KM


def _rotations(rng, n):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    ang = 2 * np.arccos(np.clip(np.abs(q[:, 0]), 0, 1))
    v = q[:, 1:] * np.sign(q[:, :1])
    return v / np.linalg.norm(v, axis=1, keepdims=True) * ang[:, None]


@pytest.mark.parametrize("case", ["uniform", "clustered", "duplicates", "quaternions"])
def test_candidate_lists_contain_the_nearest_key(case):
    rng = np.random.default_rng({"uniform": 0, "clustered": 1, "duplicates": 2, "quaternions": 3}[case])
    if case == "quaternions":
        x = rng.standard_normal((20000, 4))
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        centers, G = x[:64].copy(), 8
    else:
        x = _rotations(rng, 40000)
        if case == "clustered":
            x = x[:400][rng.integers(0, 400, 40000)] + 0.05 * rng.standard_normal((40000, 3))
        centers, G = x[:200].copy(), 16
        if case == "duplicates":
            centers[100:120] = centers[:20]                      # exact ties: the scan returns the lower id
    lo, hi = x.min(0) - 1e-9, x.max(0) + 1e-9
    blo, bhi = KM.cell_boxes(lo, hi, G)
    keep = KM.candidate_mask(centers, blo, bhi)
    cells = KM.point_cells(x, lo, hi, G)
    assert np.all((x >= blo[cells]) & (x <= bhi[cells]))          # the point -> cell map and the boxes agree
    labels = O.e_step(x, centers)                                 # sklearn's E-step (lowest index on ties)
    brute = ((x[:, None, :] - centers[None, :, :]) ** 2).sum(2).argmin(1) if case != "uniform" else labels
    assert keep[cells, labels].all(), "a nearest key is missing from its cell's candidate list"
    assert keep[cells, brute].all()
    # the pruned argmin over the candidates (ascending key order) IS the scan's answer
    d2 = ((x[:, None, :] - centers[None, :, :]) ** 2).sum(2)
    pruned = np.where(keep[cells], d2, np.inf).argmin(1)
    assert np.array_equal(pruned, d2.argmin(1))
    # and the lists are short: that is the point of the grid
    occupied = np.unique(cells)
    assert keep[occupied].sum(1).mean() < 0.25 * centers.shape[0]
