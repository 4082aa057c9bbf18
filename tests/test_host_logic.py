"""Host-side logic that needs no GPU: size functions and argument checks of the C ABI, fixed-point
scale selection, the head-stack cache rules (copies / pickles drop it, CPU parameters are rejected),
and the reference-facing error conventions."""
import copy
import pickle

import numpy as np
import pytest
import torch


def test_keygrid_sizes_and_errors():
    from bdpose import _lib
    lib = _lib.lib()
    # header + 64^3 fine records of 16 B + 64^3 / 8 side records of 64 B + fp32 copy of up to 4096 keys
    b = lib.bdp_keygrid_bytes(1000, 3)
    assert b == 160 + 64 ** 3 * 16 + (64 ** 3 // 8) * 64 + 4096 * 16
    assert lib.bdp_keygrid_bytes(16, 3) < lib.bdp_keygrid_bytes(200, 3) < b
    assert lib.bdp_keygrid_bytes(200, 4) > 0
    assert lib.bdp_keygrid_bytes(5000, 3) == -1 and lib.bdp_keygrid_bytes(100, 5) == -1
    assert lib.bdp_keygrid_build(None, 100, 3, None, 0, None) == -1
    assert b"keygrid_build" in lib.bdp_last_error()
    st = lib.bdp_assign_nearest_grid(None, _lib.F32, 10, 3, None, 100, None, 0, None, None, None, None, None)
    assert st == -1 and b"assign_nearest_grid" in lib.bdp_last_error()
    st = lib.bdp_kmeans_iteration(None, 10, 3, None, 4, None, 0, None, None, 3, None, 1, None, None, None, None)
    assert st == -1 and b"acc_stats" in lib.bdp_last_error()


def test_head_descriptor_checks_and_sizes():
    import ctypes as C
    from bdpose import _lib
    lib = _lib.lib()
    d = _lib.HeadDesc()
    d.H, d.N0, d.N1, d.N2, d.n_groups = 24, 2048, 1000, 500, 2
    d.group_heads[0] = d.group_heads[1] = 12
    d.group_out[0], d.group_out[1] = 200, 3
    B = 32
    F1, F2 = 24 * 1000, 24 * 500
    assert lib.bdp_head_saved_floats(C.byref(d), B) == 2 * B * F1 + 2 * B * F2 + 2 * F1 + 2 * F2
    assert lib.bdp_head_saved_floats(None, B) == -1
    # NULL parameter pointers are rejected before anything is launched
    assert lib.bdp_head_forward(C.byref(d), None, None, B, None, None, None) == -1
    assert b"head_forward" in lib.bdp_last_error()
    d.N1 = 1001
    assert lib.bdp_head_forward(C.byref(d), None, None, B, None, None, None) == -1
    assert b"multiples of 4" in lib.bdp_last_error()


def test_gemm_argument_checks():
    from bdpose import _lib
    lib = _lib.lib()
    assert lib.bdp_gemm_tf32(None, 0, 8, 0, None, 0, 8, 0, None, 0, 8, 0, 4, 4, 8, 1, 1, 0, 0, None) == -1
    assert lib.bdp_gemm_tf32_splits(24000, 18) == 18
    assert lib.bdp_gemm_tf32_splits(64, 18) == 2          # never an empty K split
    assert lib.bdp_gemm_tf32_splits(24000, 0) == 1


def test_fixed_point_scale():
    from bdpose import kmeans
    assert kmeans._fix_hi_bits(3.2) == 28          # |x| < 4 = 2^2 -> 2^28 * 4 = 2^30 < 2^31
    assert kmeans._fix_hi_bits(0.0) == 30
    assert kmeans._fix_hi_bits(1e12) == 0
    for m in (0.3, 1.0, 3.14159, 100.0):
        hb = kmeans._fix_hi_bits(m)
        assert m * 2.0 ** hb < 2.0 ** 31


def test_head_stack_is_a_cache():
    """deepcopy / pickle of a model drop the fused-stack cache (device buffers, ctypes descriptors)
    and CPU parameters are refused with a clear message."""
    import binDeltaModels as M
    from bdpose import head

    class FakeCuda:          # OneBinDeltaModel's ctor calls .cuda(); build the heads by hand on CPU
        pass
    bins = [M.bin_3layer.__new__(M.bin_3layer) for _ in range(2)]
    for b in bins:
        torch.nn.Module.__init__(b)
        b.fc1 = torch.nn.Linear(8, 8, bias=False); b.bn1 = torch.nn.BatchNorm1d(8)
        b.fc2 = torch.nn.Linear(8, 4, bias=False); b.bn2 = torch.nn.BatchNorm1d(4)
        b.fc3 = torch.nn.Linear(4, 5)
        object.__setattr__(b, "_solo", None)
    st = head.HeadStack([bins])
    with pytest.raises(RuntimeError, match="CUDA only"):
        st.ensure()
    assert copy.deepcopy(st) is None
    assert pickle.loads(pickle.dumps(st)) is None
    b2 = copy.deepcopy(bins[0])
    assert b2._solo is None and torch.equal(b2.fc3.weight, bins[0].fc3.weight)
    with pytest.raises(NameError):
        head.set_precision("bf16")


def test_reference_error_conventions():
    import featureModels
    with pytest.raises(NameError):
        featureModels.resnet_model("resnet18", "layer4")
    from bdpose import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.bd_loss_raw(torch.zeros(4, 8), torch.zeros(4, dtype=torch.long), None, None, None, 0, False)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under multi-modal-regression_b200/ may import it
    (only tests/, __graft_entry__.smoke() and bench.py's CPU legs do), and there is no CPU fallback
    module to route through."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "multi-modal-regression_b200")
    bad = []
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(import|from)\s+(bdpose_oracle|lloyd_host|keygrid_model|oracle)\b", src, re.M):
                    bad.append(os.path.join(d, f))
    assert not bad, "product files import the oracle: %s" % bad
