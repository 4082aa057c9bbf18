"""ctypes binding of libbdpose.so (C ABI declared in include/bdpose.h).

The library is the product; there is no fallback.  Importing this module never touches CUDA, but
`lib()` raises if the shared object has not been built (run `python -c "import __graft_entry__ as
g; g.build()"` or `make -C multi-modal-regression_b200/csrc`), and every wrapper raises
RuntimeError with the library's own message on a non-zero status.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbdpose.so")

BDP_OK = 0
# pose_mode
POSE_NONE, POSE_MSE, POSE_GEODESIC_AA, POSE_GEODESIC_Q, POSE_RIEMANNIAN, POSE_ROTMAT = range(6)
# repr / dtype
REPR_AXIS_ANGLE, REPR_QUATERNION = 0, 1
F32, F64 = 0, 1

_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_f32 = C.c_float
_f64 = C.c_double

# name -> (restype, argtypes): one entry per symbol declared in include/bdpose.h
SIGNATURES = {
    "bdp_abi_version": (_int, []),
    "bdp_last_error": (C.c_char_p, []),
    "bdp_sm_count": (_int, []),
    "bdp_bd_loss_workspace_bytes": (_i64, [_i64]),
    "bdp_bd_loss_fwd_bwd": (_int, [_p, _i64, _i64, _i64, _p, _p, _int, _p, _int, _p, _int, _p, _p,
                                   _p, _p, _p, _f32, _p, _p, _i64, _p]),
    "bdp_expected_pose_loss": (_int, [_p, _i64, _i64, _i64, _p, _int, _int, _p, _p, _int, _p, _p, _p,
                                      _p]),
    "bdp_geodesic_error_deg": (_int, [_p, _p, _int, _int, _i64, _p, _p]),
    "bdp_error_stats_workspace_bytes": (_i64, [_i64, _int]),
    "bdp_error_stats": (_int, [_p, _p, _i64, _int, _p, _p, _p, _p, _p, _i64, _p]),
    "bdp_compose_prediction": (_int, [_p, _i64, _int, _i64, _p, _int, _p, _int, _p, _p, _p]),
    "bdp_min_key_gap": (_int, [_p, _int, _int, _p, _p]),
    "bdp_sgd_step": (_int, [_p, _int, _i64, _f32, _f32, _f32, _int, _p]),
    "bdp_scale_inplace": (_int, [_p, _i64, _p, _p]),
    "bdp_fit_stats": (_int, [_p, _i64, _int, _int, _p, _f64, _p, _p, _p]),
    "bdp_assign_nearest": (_int, [_p, _int, _i64, _int, _p, _int, _p, _p, _p, _p, _p]),
    "bdp_assign_quatdot": (_int, [_p, _int, _i64, _p, _int, _p, _p, _p]),
    "bdp_assign_soft": (_int, [_p, _int, _i64, _int, _p, _int, _f64, _p, _p, _p]),
    "bdp_riemannian_residual": (_int, [_p, _int, _i64, _p, _int, _p, _p, _p, _p]),
    "bdp_convert_axis_angle": (_int, [_p, _i64, _p, _p, _p]),
    "bdp_euler_to_pose": (_int, [_p, _i64, _p, _p, _p]),
    "bdp_kmeans_lloyd_step": (_int, [_p, _i64, _int, _p, _int, _p, _p, _int, _p, _p, _int, _p]),
    "bdp_keygrid_bytes": (_i64, [_int, _int]),
    "bdp_keygrid_stats": (_int, [_p, _int, _i64, _int, _int, _p, _i64, _p, _p]),
    "bdp_keygrid_build": (_int, [_p, _int, _int, _p, _i64, _p]),
    "bdp_assign_nearest_grid": (_int, [_p, _int, _i64, _int, _p, _int, _p, _i64, _p, _p, _p, _p, _p]),
    "bdp_kmeans_lloyd_step_grid": (_int, [_p, _i64, _int, _p, _int, _p, _i64, _p, _p, _int, _p, _p,
                                          _int, _p]),
    "bdp_kmeans_iteration": (_int, [_p, _i64, _int, _p, _int, _p, _i64, _p, _p, _int, _p, _int, _p, _p,
                                    _p, _p]),
    "bdp_kmeans_finalize": (_int, [_p, _int, _int, _int, _p, _p, _p, _p, _p]),
    "bdp_kmeans_ctl_bytes": (_i64, []),
    "bdp_kmeans_xchg_bytes": (_i64, [_int, _int]),
    "bdp_kmeans_exchange_finalize": (_int, [_p, _p, _int, _int, _int, _int, _int, _int, _i64, _int, _int,
                                            _f64, _p, _p, _p, _p]),
    "bdp_kmeans_run": (_int, [_p, _i64, _int, _p, _int, _p, _i64, _p, _p, _p, _p, _int, _int, _int, _i64,
                              _int, _int, _int, _f64, _p, _p, _p, _int, _p]),
    "bdp_keygrid_coarse_cells": (_i64, [_int, _int]),
    "bdp_keygrid_prepare": (_int, [_p, _p, _int, _int, _p, _i64, _p]),
    "bdp_keygrid_occupancy": (_int, [_p, _i64, _int, _int, _p, _i64, _p, _p]),
    "bdp_cellsort_workspace_bytes": (_i64, [_i64, _int, _int]),
    "bdp_cellsort": (_int, [_p, _i64, _int, _int, _p, _i64, _p, _p, _i64, _p, _p, _p]),
    "bdp_scatter_i32": (_int, [_p, _p, _i64, _p, _p]),
    "bdp_gemm_tf32": (_int, [_p, _int, _i64, _i64, _p, _int, _i64, _i64, _p, _int, _i64, _i64, _i64,
                             _i64, _i64, _int, _int, _i64, _int, _p]),
    "bdp_gemm_tf32_splits": (_int, [_i64, _int]),
    "bdp_bn_relu_fwd": (_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _f32, _f32, _int, _p,
                               _p]),
    "bdp_bn_relu_bwd": (_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _int, _p, _p, _p, _p]),
    "bdp_head_fc3_fwd": (_int, [_p, _i64, _p, _p, _p, _i64, _int, _int, _int, _p, _p]),
    "bdp_head_fc3_bwd": (_int, [_p, _p, _i64, _p, _p, _p, _i64, _int, _int, _int, _p, _p, _p, _p,
                                _p]),
    "bdp_sum_slabs": (_int, [_p, _i64, _int, _i64, _p, _p]),
    "bdp_head_saved_floats": (_i64, [_p, _i64]),
    "bdp_head_bwd_workspace_floats": (_i64, [_p, _i64]),
    "bdp_head_forward": (_int, [_p, _p, _p, _i64, _p, _p, _p]),
    "bdp_head_backward": (_int, [_p, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                 _p]),
}

HEAD_MAX_GROUPS = 4
COMPOSE_ADD, COMPOSE_ADD_NORMALIZE, COMPOSE_RIEMANNIAN = range(3)


class SgdTensor(C.Structure):
    """struct bdp_sgd_tensor (include/bdpose.h)"""
    _fields_ = [("p", _p), ("g", _p), ("buf", _p), ("n", C.c_int64), ("step_size", C.c_float),
                ("first", C.c_int32)]

KMEANS_MAX_RANKS = 8
KMEANS_RUNNING, KMEANS_STRICT, KMEANS_TOL, KMEANS_NEEDS_HOST = range(4)


class KMeansStatus(C.Structure):
    """struct bdp_kmeans_status (include/bdpose.h): head of the device control block"""
    _fields_ = [("state", C.c_int32), ("reserved", C.c_int32), ("iter_done", C.c_int64),
                ("changed", C.c_int64), ("n_empty", C.c_int64), ("shift2", C.c_double)]



class HeadDesc(C.Structure):
    """struct bdp_head_desc (include/bdpose.h)"""
    _fields_ = [("H", C.c_int32), ("N0", C.c_int32), ("N1", C.c_int32), ("N2", C.c_int32),
                ("n_groups", C.c_int32), ("training", C.c_int32), ("precise", C.c_int32),
                ("reserved", C.c_int32),
                ("group_heads", C.c_int32 * HEAD_MAX_GROUPS), ("group_out", C.c_int32 * HEAD_MAX_GROUPS),
                ("w1", _p), ("g1", _p), ("be1", _p), ("w2", _p), ("g2", _p), ("be2", _p),
                ("rm1", _p), ("rv1", _p), ("rm2", _p), ("rv2", _p),
                ("w3", _p * HEAD_MAX_GROUPS), ("b3", _p * HEAD_MAX_GROUPS),
                ("eps", C.c_float), ("momentum", C.c_float)]

_lib = None


def lib():
    """The loaded shared library (loaded once).  Raises if it is missing: no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libbdpose.so is not built (%s). Build it with `python -c 'import __graft_entry__ as "
                "g; g.build()'`; this package has no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status, what):
    if status != BDP_OK:
        msg = lib().bdp_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, status, msg.decode() if msg else "?"))


def ptr(t):
    """Device (or host) address of a tensor's first element, or NULL for None."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
