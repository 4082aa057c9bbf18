"""Tensor-level wrappers over the C ABI (include/bdpose.h).

Every function takes CUDA tensors, allocates its outputs with torch (the caller of the C ABI owns
all memory), launches on torch's current stream and never synchronises unless it has to return a
Python number.  CPU tensors are rejected: there is no CPU path in this package.
"""
import torch

from . import _lib as L

_loss_ws = {}     # (device index, stream) -> zero-initialised workspace, reused (kernel re-zeroes it)
_stats_ws = {}


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("bdpose: expected CUDA tensors (this package has no CPU path); got a "
                               "%s tensor" % t.device)


def _loss_workspace(dev):
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _loss_ws.get(key)
    if ws is None:
        n = L.lib().bdp_bd_loss_workspace_bytes(1)
        ws = torch.zeros(n, dtype=torch.uint8, device=dev)
        _loss_ws[key] = ws
    return ws


def _dtype_code(t):
    if t.dtype == torch.float32:
        return L.F32
    if t.dtype == torch.float64:
        return L.F64
    raise TypeError("bdpose: expected float32 or float64, got %s" % t.dtype)


# ------------------------------------------------------------------------------------------------
# (b) fused bin-delta loss
# ------------------------------------------------------------------------------------------------
def bd_loss_raw(logits, bin_true, pred, target, keys, pose_mode, use_keys, want_grad=True,
                per_sample=False, want_rows=False):
    """One launch of bdp_bd_loss_fwd_bwd.

    Returns dict(loss=[2] fp32 {mean CE, mean pose}, grad_logits, grad_pred, argmax, row_ce,
    row_pose).  Gradients are those of the MEAN losses (scale 1/B) unless per_sample.
    """
    _need_cuda(logits, bin_true, pred, target, keys)
    ref = logits if logits is not None else pred
    dev = ref.device
    B = ref.shape[0]
    K = ld = 0
    if logits is not None:
        if logits.dtype != torch.float32:
            logits = logits.float()
        if logits.stride(1) != 1 or logits.dim() != 2:
            logits = logits.contiguous()
        K, ld = logits.shape[1], logits.stride(0)
        if ld < K:
            logits = logits.contiguous()
            ld = K
        bin_true = bin_true.reshape(-1).to(torch.int64).contiguous()
        if bin_true.numel() != B:
            raise ValueError("bd_loss: bin target has %d entries for %d rows" % (bin_true.numel(), B))
    ndim = 0
    if pose_mode != L.POSE_NONE:
        pred = pred.float().reshape(B, -1).contiguous()
        target = target.float().reshape(B, -1).contiguous()
        ndim = pred.shape[1]
        want_t = 9 if pose_mode in (L.POSE_RIEMANNIAN, L.POSE_ROTMAT) else ndim
        if target.shape[1] != want_t:
            raise ValueError("bd_loss: target has %d columns, expected %d" % (target.shape[1], want_t))
        if keys is not None:
            keys = keys.float().reshape(keys.shape[0], -1).contiguous()
            if logits is not None and keys.shape[0] != K:
                raise ValueError("bd_loss: %d keys for %d bins" % (keys.shape[0], K))
    out = torch.empty(2, dtype=torch.float32, device=dev)
    g_logits = None
    if want_grad and logits is not None:
        g_logits = torch.empty((B, ld), dtype=torch.float32, device=dev)
    g_pred = torch.empty((B, ndim), dtype=torch.float32, device=dev) \
        if (want_grad and pose_mode != L.POSE_NONE) else None
    amax = torch.empty(B, dtype=torch.int64, device=dev) if logits is not None else None
    row_ce = torch.empty(B, dtype=torch.float32, device=dev) \
        if (want_rows and logits is not None) else None
    row_pose = torch.empty(B, dtype=torch.float32, device=dev) \
        if ((want_rows or per_sample) and pose_mode != L.POSE_NONE) else None
    ws = _loss_workspace(dev)
    with torch.cuda.device(dev):
        st = L.lib().bdp_bd_loss_fwd_bwd(
            L.ptr(logits), B, K, ld, L.ptr(bin_true), L.ptr(pred), ndim, L.ptr(keys),
            1 if use_keys else 0, L.ptr(target), pose_mode, L.ptr(out), L.ptr(row_ce),
            L.ptr(row_pose), L.ptr(g_logits), L.ptr(g_pred), 1.0 if per_sample else 0.0,
            L.ptr(amax), L.ptr(ws), ws.numel(), L.stream_ptr())
    L.check(st, "bdp_bd_loss_fwd_bwd")
    if g_logits is not None and ld != K:
        g_logits = g_logits[:, :K]
    return dict(loss=out, grad_logits=g_logits, grad_pred=g_pred, argmax=amax, row_ce=row_ce,
                row_pose=row_pose)


def _scale_inplace(buf, g):
    """buf *= g (0-dim tensor on the device) without a host read of g."""
    if not (buf.is_contiguous() and buf.data_ptr() % 16 == 0):
        return buf * g
    g = g.detach().to(torch.float32).reshape(1)
    with torch.cuda.device(buf.device):
        st = L.lib().bdp_scale_inplace(L.ptr(buf), buf.numel(), L.ptr(g), L.stream_ptr())
    L.check(st, "bdp_scale_inplace")
    return buf


class _BDLoss(torch.autograd.Function):
    """(Lc, Lr) = fused(logits, pred); the backward only scales the gradients the forward launch
    already wrote (saved per call, so two forwards before one backward are fine —
    learnGeodesicBDModel.py:116-120,183)."""

    @staticmethod
    def forward(ctx, logits, pred, bin_true, target, keys, pose_mode, use_keys):
        need = [logits is not None and logits.requires_grad, pred is not None and pred.requires_grad]
        r = bd_loss_raw(logits, bin_true, pred, target, keys, pose_mode, use_keys,
                        want_grad=any(need))
        ctx.save_for_backward(r["grad_logits"], r["grad_pred"])
        ctx.shapes = (None if logits is None else logits.shape, None if pred is None else pred.shape)
        ctx.scaled = False
        ctx.mark_non_differentiable(r["argmax"]) if r["argmax"] is not None else None
        lc, lr = r["loss"][0], r["loss"][1]
        if r["argmax"] is None:
            return lc, lr
        return lc, lr, r["argmax"]

    @staticmethod
    def backward(ctx, g_lc, g_lr, *unused):
        g_logits, g_pred = ctx.saved_tensors
        s_logits, s_pred = ctx.shapes
        gl = gp = None
        if g_logits is not None and ctx.needs_input_grad[0]:
            # The forward launch already wrote d Lc / d logits; the upstream scalar (1 for
            # `Lc + w * Lr`) is applied in place by a kernel that returns without touching memory
            # when it is 1 — no second pass over [B, K].
            if ctx.scaled:
                raise RuntimeError("bd_loss: backward through the same graph twice (the saved "
                                   "gradient was scaled in place); re-run the forward")
            ctx.scaled = True
            gl = _scale_inplace(g_logits, g_lc).reshape(s_logits)
        if g_pred is not None and ctx.needs_input_grad[1]:
            gp = (g_pred * g_lr).reshape(s_pred)
        return gl, gp, None, None, None, None, None


def bd_loss(logits, bin_true, pred, target, keys=None, pose_mode=L.POSE_NONE, use_keys=False):
    """Fused CE + pose loss with autograd.  Returns (Lc, Lr, argmax_bin) as 0-dim / [B] tensors."""
    out = _BDLoss.apply(logits, pred, bin_true, target, keys, pose_mode, use_keys)
    return out


class _PoseLossRows(torch.autograd.Function):
    """reduce=False pose loss: per-sample values, per-sample upstream gradients.  The gradient w.r.t.
    the SECOND argument is produced too when it is asked for: the soft-bin losses call
    `my_loss(ydata, centers[k] + residual)` with the prediction in the `ytrue` slot
    (binDeltaLosses.py:120, 143, 164)."""

    @staticmethod
    def forward(ctx, pred, target, pose_mode):
        r = bd_loss_raw(None, None, pred, target, None, pose_mode, False,
                        want_grad=pred.requires_grad, per_sample=True)
        g_target = None
        if target.requires_grad:
            if pose_mode == L.POSE_GEODESIC_AA:
                # the axis-angle geodesic distance is symmetric in its arguments (both are normalised,
                # axisAngle.py:110-120): d/d target = d/d pred of the swapped call
                g_target = bd_loss_raw(None, None, target, pred, None, pose_mode, False,
                                       want_grad=True, per_sample=True)["grad_pred"]
            else:
                raise RuntimeError("pose loss: gradient w.r.t. the second argument is only fused for "
                                   "the axis-angle geodesic loss")
        ctx.save_for_backward(r["grad_pred"], g_target)
        ctx.shapes = (pred.shape, target.shape)
        return r["row_pose"]

    @staticmethod
    def backward(ctx, g_rows):
        g_pred, g_target = ctx.saved_tensors
        gp = gt = None
        if g_pred is not None and ctx.needs_input_grad[0]:
            gp = (g_pred * g_rows.reshape(-1, 1)).reshape(ctx.shapes[0])
        if g_target is not None and ctx.needs_input_grad[1]:
            gt = (g_target * g_rows.reshape(-1, 1)).reshape(ctx.shapes[1])
        return gp, gt, None


def _quat_geodesic_torch(pred, target, reduce):
    """quaternion.geodesic_loss (quaternion.py:156-163) in torch ops: used when the gradient has to
    reach the un-normalised second argument (the loss is not symmetric in its arguments)."""
    qh = torch.nn.functional.normalize(pred.float(), dim=1)
    w = torch.sum(target.float() * qh, dim=1)
    theta = 2.0 * torch.acos(torch.clamp(torch.abs(w), -1.0 + 1e-6, 1.0 - 1e-6))
    return theta.mean() if reduce else theta


def pose_loss(pred, target, pose_mode, reduce=True):
    """Stand-alone pose loss (axisAngle.geodesic_loss, quaternion.geodesic_loss, rotmat my_loss)."""
    two_sided = torch.is_grad_enabled() and isinstance(target, torch.Tensor) and target.requires_grad
    if two_sided:
        if pose_mode == L.POSE_GEODESIC_Q:
            return _quat_geodesic_torch(pred, target, reduce)
        if pose_mode == L.POSE_GEODESIC_AA:
            rows = _PoseLossRows.apply(pred, target, pose_mode)
            return rows.mean() if reduce else rows
        raise RuntimeError("pose loss: the second argument requires a gradient, which mode %d does "
                           "not provide" % pose_mode)
    if reduce:
        return _BDLoss.apply(None, pred, None, target, None, pose_mode, False)[1]
    return _PoseLossRows.apply(pred, target, pose_mode)


class _ExpectedPose(torch.autograd.Function):
    """rows[b] = sum_k softmax(score_b)_k * L(target_b, keys_k + delta_b[k]) in one launch."""

    @staticmethod
    def forward(ctx, score, delta, target, keys, pose_mode, per_bin):
        B, K = score.shape
        nd = keys.shape[1]
        score_c = score.detach().float().contiguous()
        delta_c = delta.detach().float().contiguous()
        rows = torch.empty(B, dtype=torch.float32, device=score.device)
        g_s = torch.empty((B, K), dtype=torch.float32, device=score.device)
        g_d = torch.empty_like(delta_c)
        with torch.cuda.device(score.device):
            st = L.lib().bdp_expected_pose_loss(
                L.ptr(score_c), B, K, K, L.ptr(delta_c), 1 if per_bin else 0, nd, L.ptr(keys),
                L.ptr(target), pose_mode, L.ptr(rows), L.ptr(g_s), L.ptr(g_d), L.stream_ptr())
        L.check(st, "bdp_expected_pose_loss")
        ctx.save_for_backward(g_s, g_d)
        ctx.dshape = delta.shape
        return rows

    @staticmethod
    def backward(ctx, g_rows):
        g_s, g_d = ctx.saved_tensors
        gs = g_s * g_rows.reshape(-1, 1) if ctx.needs_input_grad[0] else None
        gd = None
        if ctx.needs_input_grad[1]:
            shape = [-1] + [1] * (g_d.dim() - 1)
            gd = (g_d * g_rows.reshape(shape)).reshape(ctx.dshape)
        return gs, gd, None, None, None, None


def expected_pose_loss(score, delta, target, keys, pose_mode):
    """Per-row expectation of the pose loss over the bins (soft-bin losses).  score [B,K],
    delta [B,nd] or [B,K,nd], target [B,nd], keys [K,nd]; returns [B] with autograd to score, delta."""
    _need_cuda(score, delta, target, keys)
    K = score.shape[1]
    keys = keys.detach().float().reshape(K, -1).contiguous()
    target = target.detach().float().reshape(score.shape[0], -1).contiguous()
    per_bin = delta.dim() == 3
    return _ExpectedPose.apply(score, delta, target, keys, pose_mode, per_bin)


# ------------------------------------------------------------------------------------------------
# (d) evaluation
# ------------------------------------------------------------------------------------------------
def geodesic_error_deg(y_gt, y_hat, quaternion=False):
    """[N] fp64 errors in degrees (axisAngle.get_error / quaternion.get_error, per-sample part)."""
    _need_cuda(y_gt, y_hat)
    d = 4 if quaternion else 3
    if y_gt.dtype != y_hat.dtype or y_gt.dtype not in (torch.float32, torch.float64):
        y_gt, y_hat = y_gt.double(), y_hat.double()
    y_gt = y_gt.reshape(-1, d).contiguous()
    y_hat = y_hat.reshape(-1, d).contiguous()
    if y_gt.shape != y_hat.shape:
        raise ValueError("geodesic_error: shape mismatch %s vs %s" % (y_gt.shape, y_hat.shape))
    N = y_gt.shape[0]
    err = torch.empty(N, dtype=torch.float64, device=y_gt.device)
    with torch.cuda.device(y_gt.device):
        st = L.lib().bdp_geodesic_error_deg(
            L.ptr(y_gt), L.ptr(y_hat), _dtype_code(y_gt),
            L.REPR_QUATERNION if quaternion else L.REPR_AXIS_ANGLE, N, L.ptr(err), L.stream_ptr())
    L.check(st, "bdp_geodesic_error_deg")
    return err


def error_stats(err_deg, labels=None, num_classes=1):
    """(median[C] fp64, count[C] int64, below30[1] int64, max[1] fp64) on the device."""
    _need_cuda(err_deg, labels)
    err_deg = err_deg.double().reshape(-1).contiguous()
    N = err_deg.numel()
    dev = err_deg.device
    if labels is not None:
        labels = labels.reshape(-1).to(torch.int64).contiguous()
        if labels.numel() != N:
            raise ValueError("error_stats: %d labels for %d errors" % (labels.numel(), N))
    C = int(num_classes)
    med = torch.empty(C, dtype=torch.float64, device=dev)
    cnt = torch.empty(C, dtype=torch.int64, device=dev)
    b30 = torch.empty(1, dtype=torch.int64, device=dev)
    mx = torch.empty(1, dtype=torch.float64, device=dev)
    n = L.lib().bdp_error_stats_workspace_bytes(N, C)
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, n)
    ws = _stats_ws.get(key)
    if ws is None:
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        _stats_ws[key] = ws
    with torch.cuda.device(dev):
        st = L.lib().bdp_error_stats(L.ptr(err_deg), L.ptr(labels), N, C, L.ptr(med), L.ptr(cnt),
                                     L.ptr(b30), L.ptr(mx), L.ptr(ws), n, L.stream_ptr())
    L.check(st, "bdp_error_stats")
    return med, cnt, b30, mx


# ------------------------------------------------------------------------------------------------
# (c) assignment
# ------------------------------------------------------------------------------------------------
# Nearest-key queries go through the key grid (candidate pruning, include/bdpose.h) when the
# dictionary and the batch are big enough for the two small build launches to pay off.
GRID_MIN_K = 8
GRID_MIN_N = 1 << 15
_grid_ws = {}     # (device index, stream, bytes) -> grid buffer, reused (rebuilt on every call)


class KeyGrid:
    """Device buffer holding the key grid of one dictionary (bdp_keygrid_build)."""

    def __init__(self, centers, buf=None, build=True):
        _need_cuda(centers)
        self.centers = centers.double().contiguous()
        self.K, self.d = self.centers.shape
        self.nbytes = L.lib().bdp_keygrid_bytes(self.K, self.d)
        if self.nbytes < 0:
            raise RuntimeError("key grid: unsupported dictionary shape [%d, %d]" % (self.K, self.d))
        dev = self.centers.device
        if buf is None or buf.numel() < self.nbytes or buf.device != dev:
            buf = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)
        self.buf = buf
        if build:                 # build=False: the owner builds it (sharded multi-GPU build)
            self.rebuild()

    def rebuild(self, centers=None):
        """(Re)build for the current / new centre values (same shape); no host synchronisation."""
        if centers is not None:
            self.centers = centers.double().contiguous()
        with torch.cuda.device(self.buf.device):
            st = L.lib().bdp_keygrid_build(L.ptr(self.centers), self.K, self.d, L.ptr(self.buf),
                                           self.nbytes, L.stream_ptr())
        L.check(st, "bdp_keygrid_build")
        return self

    @staticmethod
    def supported(K, d, N):
        return d in (3, 4) and GRID_MIN_K <= K <= 4096 and N >= GRID_MIN_N


def keygrid_stats(grid, x):
    """Candidate-list statistics of the rotations x [N,d] against a built KeyGrid (diagnostics)."""
    _need_cuda(x)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.double()
    x = x.reshape(x.shape[0], -1).contiguous()
    st = torch.zeros(5, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.lib().bdp_keygrid_stats(L.ptr(x), _dtype_code(x), x.shape[0], x.shape[1], grid.K,
                                       L.ptr(grid.buf), grid.nbytes, L.ptr(st), L.stream_ptr())
    L.check(rc, "bdp_keygrid_stats")
    n, outside, over, length, mx = [int(v) for v in st.tolist()]
    fast = max(n - outside - over, 1)
    return {"points": n, "slow_path_fraction": (outside + over) / max(n, 1),
            "outside_grid": outside, "overflow_cells": over, "mean_list_length": length / fast,
            "max_list_length": mx}


def _scratch_grid(centers):
    dev = centers.device
    nbytes = L.lib().bdp_keygrid_bytes(centers.shape[0], centers.shape[1])
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _grid_ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _grid_ws[key] = buf
    return KeyGrid(centers, buf)


def assign_nearest(x, centers, want_residual=True, want_sqdist=False, label_dtype=torch.int64,
                   grid="auto", out_labels=None, out_residual=None):
    """kmeans.predict + residual (binDeltaGenerators.py:27-30).  x [N,d] fp32|fp64, centers [K,d].

    grid: "auto" (build a key grid when it pays off), None (brute-force scan) or a KeyGrid built for
    exactly these centres.  out_labels / out_residual: preallocated contiguous CUDA outputs (the
    pipelined host-to-host path reuses its buffers).  Returns (labels [N] label_dtype,
    residual [N,d] fp32 | None, sqdist [N] fp64 | None)."""
    _need_cuda(x, centers)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.double()
    x = x.reshape(x.shape[0], -1).contiguous()
    centers = centers.double().contiguous()
    N, d = x.shape
    K = centers.shape[0]
    if centers.shape[1] != d:
        raise ValueError("assign_nearest: x has %d columns, centers %d" % (d, centers.shape[1]))
    dev = x.device
    if out_labels is not None:
        if out_labels.dtype != label_dtype or out_labels.numel() != N or not out_labels.is_contiguous() \
                or out_labels.device != dev:
            raise ValueError("assign_nearest: out_labels must be a contiguous %s [%d] tensor on %s" % (label_dtype, N, dev))
    if out_residual is not None:
        if out_residual.dtype != torch.float32 or tuple(out_residual.shape) != (N, d) or \
                not out_residual.is_contiguous() or out_residual.device != dev:
            raise ValueError("assign_nearest: out_residual must be a contiguous float32 [%d, %d] tensor on %s" % (N, d, dev))
    lab32 = lab64 = None
    if label_dtype == torch.int32:
        lab32 = out_labels if out_labels is not None else torch.empty(N, dtype=torch.int32, device=dev)
    else:
        lab64 = out_labels if out_labels is not None else torch.empty(N, dtype=torch.int64, device=dev)
    res = None
    if want_residual:
        res = out_residual if out_residual is not None else torch.empty((N, d), dtype=torch.float32, device=dev)
    sq = torch.empty(N, dtype=torch.float64, device=dev) if want_sqdist else None
    if isinstance(grid, str):
        if grid != "auto":
            raise NameError("Unknown grid mode passed")
        grid = _scratch_grid(centers) if KeyGrid.supported(K, d, N) else None
    with torch.cuda.device(dev):
        if grid is None:
            st = L.lib().bdp_assign_nearest(L.ptr(x), _dtype_code(x), N, d, L.ptr(centers), K,
                                            L.ptr(lab32), L.ptr(lab64), L.ptr(res), L.ptr(sq),
                                            L.stream_ptr())
        else:
            if grid.K != K or grid.d != d:
                raise ValueError("assign_nearest: key grid was built for a [%d, %d] dictionary" % (grid.K, grid.d))
            st = L.lib().bdp_assign_nearest_grid(L.ptr(x), _dtype_code(x), N, d, L.ptr(centers), K,
                                                 L.ptr(grid.buf), grid.nbytes, L.ptr(lab32),
                                                 L.ptr(lab64), L.ptr(res), L.ptr(sq), L.stream_ptr())
    L.check(st, "bdp_assign_nearest")
    return (lab32 if lab32 is not None else lab64), res, sq


def assign_quatdot(q, keys):
    """bin = argmax_k |<key_k, q>|, residual = q - keys[bin] (learnObjectnetModel.py:108-109)."""
    _need_cuda(q, keys)
    if q.dtype not in (torch.float32, torch.float64):
        q = q.double()
    q = q.reshape(-1, 4).contiguous()
    keys = keys.double().reshape(-1, 4).contiguous()
    N = q.shape[0]
    b = torch.empty(N, dtype=torch.int64, device=q.device)
    res = torch.empty((N, 4), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        st = L.lib().bdp_assign_quatdot(L.ptr(q), _dtype_code(q), N, L.ptr(keys), keys.shape[0],
                                        L.ptr(b), L.ptr(res), L.stream_ptr())
    L.check(st, "bdp_assign_quatdot")
    return b, res


def assign_soft(x, centers, gamma=10.0, want_p=True, want_residual=True):
    """p = softmax_k(-gamma ||x - c_k||^2) (exp / normalise in fp64), residual = x - p @ centers
    (binDeltaGenerators.py:104-108).  x [N,d] fp32|fp64, d = 3|4 -> (p [N,K] fp32, residual [N,d] fp32)."""
    _need_cuda(x, centers)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.double()
    centers = centers.double().contiguous()
    K, d = centers.shape
    x = x.reshape(-1, d).contiguous()
    N = x.shape[0]
    p = torch.empty((N, K), dtype=torch.float32, device=x.device) if want_p else None
    res = torch.empty((N, d), dtype=torch.float32, device=x.device) if want_residual else None
    with torch.cuda.device(x.device):
        st = L.lib().bdp_assign_soft(L.ptr(x), _dtype_code(x), N, d, L.ptr(centers), K, float(gamma),
                                     L.ptr(p), L.ptr(res), L.stream_ptr())
    L.check(st, "bdp_assign_soft")
    return p, res


def euler_to_pose(euler_deg, want_aa=True, want_quat=False):
    """(az, el, ct) in degrees [N,3] -> axis-angle [N,3] and / or quaternion [N,4], fp64
    (helperFunctions.rotation_matrix + axisAngle.get_y / quaternion.get_y)."""
    _need_cuda(euler_deg)
    e = euler_deg.double().reshape(-1, 3).contiguous()
    N = e.shape[0]
    aa = torch.empty((N, 3), dtype=torch.float64, device=e.device) if want_aa else None
    q = torch.empty((N, 4), dtype=torch.float64, device=e.device) if want_quat else None
    with torch.cuda.device(e.device):
        st = L.lib().bdp_euler_to_pose(L.ptr(e), N, L.ptr(aa), L.ptr(q), L.stream_ptr())
    L.check(st, "bdp_euler_to_pose")
    return aa, q


def riemannian_residual(x, key_rot=None, bins=None, want_rot=True):
    """ydata_rot = get_R(x) and res = get_y(R_key[bin]^T R) (binDeltaGenerators.py:131-137)."""
    _need_cuda(x, key_rot, bins)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.double()
    x = x.reshape(-1, 3).contiguous()
    N = x.shape[0]
    dev = x.device
    rot = torch.empty((N, 3, 3), dtype=torch.float32, device=dev) if want_rot else None
    res = None
    K = 0
    if key_rot is not None:
        key_rot = key_rot.double().reshape(-1, 9).contiguous()
        K = key_rot.shape[0]
        bins = bins.reshape(-1).to(torch.int64).contiguous()
        res = torch.empty((N, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = L.lib().bdp_riemannian_residual(L.ptr(x), _dtype_code(x), N, L.ptr(key_rot), K,
                                             L.ptr(bins), L.ptr(rot), L.ptr(res), L.stream_ptr())
    L.check(st, "bdp_riemannian_residual")
    return rot, res


def convert_axis_angle(aa, want_rot=True, want_quat=True):
    """axis-angle [N,3] -> (rotation matrices [N,3,3] fp64, unit quaternions [N,4] fp64)."""
    _need_cuda(aa)
    aa = aa.double().reshape(-1, 3).contiguous()
    N = aa.shape[0]
    rot = torch.empty((N, 3, 3), dtype=torch.float64, device=aa.device) if want_rot else None
    quat = torch.empty((N, 4), dtype=torch.float64, device=aa.device) if want_quat else None
    with torch.cuda.device(aa.device):
        st = L.lib().bdp_convert_axis_angle(L.ptr(aa), N, L.ptr(rot), L.ptr(quat), L.stream_ptr())
    L.check(st, "bdp_convert_axis_angle")
    return rot, quat


# ------------------------------------------------------------------------------------------------
# (d3) test-time pose composition, (f2) get_gamma
# ------------------------------------------------------------------------------------------------
def compose_prediction(score, residual, dictionary, mode="add"):
    """ypred of the scripts' testing() loops on the device: bin = argmax(score), then
    mode "add": dict[bin] + res (learnGeodesicBDModel.py:217-219); "normalize": that, divided by
    max(norm, 1e-10) (learnGeodesicBDModel_quaternion.py:217-218); "riemannian": dictionary = key
    rotations [K,3,3], get_y(R_key[bin] . get_R(res)) (learnRiemannianBDModel.py:247).
    Returns (ypred [N, ndim] fp64, bin [N] int64)."""
    _need_cuda(score, residual, dictionary)
    modes = {"add": L.COMPOSE_ADD, "normalize": L.COMPOSE_ADD_NORMALIZE, "riemannian": L.COMPOSE_RIEMANNIAN}
    if mode not in modes:
        raise NameError("Unknown composition mode passed")
    score = score.float()
    if score.dim() != 2 or score.stride(1) != 1:
        score = score.contiguous()
    N, K = score.shape
    residual = residual.float().reshape(N, -1).contiguous()
    nd = residual.shape[1]
    dictionary = dictionary.double().reshape(K, -1).contiguous()
    want = 9 if mode == "riemannian" else nd
    if dictionary.shape[1] != want:
        raise ValueError("compose_prediction: dictionary has %d columns, expected %d" % (dictionary.shape[1], want))
    out = torch.empty((N, nd), dtype=torch.float64, device=score.device)
    bins = torch.empty(N, dtype=torch.int64, device=score.device)
    with torch.cuda.device(score.device):
        st = L.lib().bdp_compose_prediction(L.ptr(score), N, K, score.stride(0), L.ptr(residual), nd,
                                            L.ptr(dictionary), modes[mode], L.ptr(out), L.ptr(bins),
                                            L.stream_ptr())
    L.check(st, "bdp_compose_prediction")
    return out, bins


def min_key_gap(keys):
    """min_{i != j} ||k_i - k_j||^2 of a dictionary [K, d] (0-dim fp64 tensor on the device)."""
    _need_cuda(keys)
    keys = keys.double().contiguous()
    out = torch.empty(1, dtype=torch.float64, device=keys.device)
    with torch.cuda.device(keys.device):
        st = L.lib().bdp_min_key_gap(L.ptr(keys), keys.shape[0], keys.shape[1], L.ptr(out), L.stream_ptr())
    L.check(st, "bdp_min_key_gap")
    return out[0]
