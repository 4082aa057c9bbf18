"""The C-ABI library loads and exports every symbol include/bdpose.h declares (CPU: no compute)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bdpose.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bdp_\w+)\s*\(", src)))


def test_build_and_symbols():
    import __graft_entry__ as g
    g.build()
    from bdpose import _lib
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 14
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libbdpose.so does not export %s" % n
    # the Python binding covers exactly the declared surface
    assert sorted(_lib.SIGNATURES) == names
    assert lib.bdp_abi_version() == 1


def test_argument_errors_without_gpu():
    """Bad arguments are rejected before any CUDA call, with a message."""
    from bdpose import _lib
    lib = _lib.lib()
    st = lib.bdp_assign_nearest(None, _lib.F32, 10, 5, None, 4, None, None, None, None, None)
    assert st == -1
    assert b"assign_nearest" in lib.bdp_last_error()
    st = lib.bdp_bd_loss_fwd_bwd(None, 0, 0, 0, None, None, 0, None, 0, None, 0, None, None, None,
                                 None, None, 0.0, None, None, 0, None)
    assert st == -1
    assert lib.bdp_assign_nearest(None, _lib.F32, 0, 3, None, 4, None, None, None, None, None) == 0


def test_no_cpu_path():
    import pytest
    import torch
    from bdpose import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.assign_nearest(torch.zeros(4, 3), torch.zeros(2, 3, dtype=torch.float64))
