"""The category-conditioned bin-delta head on the B200 kernels.

`HeadStack` owns the STACKED parameters of H identical 3-layer MLPs (fc1 [H,N1,N0], bn1 [H,N1],
fc2 [H,N2,N1], bn2 [H,N2], fc3 [H,O,N2] + bias) and runs all of them on one feature batch:

  fc1   one tcgen05 TF32 GEMM over the flattened [H*N1, N0] weights        (bdp_gemm_tf32)
  bn1   batch statistics + affine + ReLU on feature-major activations       (bdp_bn_relu_fwd)
  fc2   grouped tcgen05 GEMM, one [N2,N1] weight per head                   (bdp_gemm_tf32, G=H)
  bn2   as bn1
  fc3   label-selected / soft-mixed output layer                            (bdp_head_fc3_fwd)

and the hand-derived backward (dgrad / wgrad through the same GEMM kernel with MN-major operand
descriptors, BatchNorm backward, fc3 backward).  Every head sees every sample in train mode, exactly
as the reference does (binDeltaModels.py:114-115): BatchNorm couples the batch inside each head, so
routing samples to their own category's head only would change the statistics.

Precision: the GEMMs run on the tensor cores in TF32 (10-bit mantissa operands, fp32 accumulate)
straight from the fp32 master weights — the "reduced-precision head GEMM" of the north star, with a
tighter error than bf16 (tolerance 2e-3 relative in the tests).
"""
import torch

from . import _lib as L

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _pad4(n):
    return (n + 3) // 4 * 4


def gemm_tf32(A, a_major, a_ld, a_gs, B, b_major, b_ld, b_gs, C, c_layout, ldc, c_gs, M, N, K, G=1,
              splits=1, c_ss=0):
    """Raw launch of bdp_gemm_tf32 (see include/bdpose.h for the operand conventions)."""
    with torch.cuda.device(A.device):
        st = L.lib().bdp_gemm_tf32(L.ptr(A), a_major, a_ld, a_gs, L.ptr(B), b_major, b_ld, b_gs,
                                   L.ptr(C), c_layout, ldc, c_gs, M, N, K, G, splits, c_ss,
                                   L.stream_ptr())
    L.check(st, "bdp_gemm_tf32")


def gemm_splits(K, splits):
    return L.lib().bdp_gemm_tf32_splits(K, splits)


def bn_relu_fwd(h, B, gamma, beta, running_mean, running_var, training, eps=BN_EPS,
                momentum=BN_MOMENTUM):
    """h [F, ldb] -> (a [F, ldb], save_mean [F], save_invstd [F])."""
    F, ldb = h.shape
    a = torch.empty_like(h)
    if training:
        mean = torch.empty(F, dtype=torch.float32, device=h.device)
        invstd = torch.empty(F, dtype=torch.float32, device=h.device)
    else:
        mean = invstd = None
    with torch.cuda.device(h.device):
        st = L.lib().bdp_bn_relu_fwd(L.ptr(h), F, B, ldb, L.ptr(gamma), L.ptr(beta),
                                     L.ptr(running_mean), L.ptr(running_var), L.ptr(mean),
                                     L.ptr(invstd), eps, momentum, 1 if training else 0, L.ptr(a),
                                     L.stream_ptr())
    L.check(st, "bdp_bn_relu_fwd")
    return a, mean, invstd


def bn_relu_bwd(da, a, h, gamma, mean, invstd, B, training=True):
    F, ldb = h.shape
    dh = torch.empty_like(h)
    dgamma = torch.empty(F, dtype=torch.float32, device=h.device)
    dbeta = torch.empty(F, dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        st = L.lib().bdp_bn_relu_bwd(L.ptr(da), L.ptr(a), L.ptr(h), L.ptr(gamma), L.ptr(mean),
                                     L.ptr(invstd), F, B, ldb, 1 if training else 0, L.ptr(dh),
                                     L.ptr(dgamma), L.ptr(dbeta), L.stream_ptr())
    L.check(st, "bdp_bn_relu_bwd")
    return dh, dgamma, dbeta


def fc3_fwd(a2, w3, b3, mix, B):
    H, O, N2 = w3.shape
    y = torch.empty((B, O), dtype=torch.float32, device=a2.device)
    with torch.cuda.device(a2.device):
        st = L.lib().bdp_head_fc3_fwd(L.ptr(a2), a2.shape[1], L.ptr(w3), L.ptr(b3), L.ptr(mix), B,
                                      H, O, N2, L.ptr(y), L.stream_ptr())
    L.check(st, "bdp_head_fc3_fwd")
    return y


def fc3_bwd(dy, a2, w3, b3, mix, B, want_dmix):
    H, O, N2 = w3.shape
    da2 = torch.empty_like(a2)
    dw3 = torch.empty_like(w3)
    db3 = torch.empty_like(b3)
    dmix = torch.empty((B, H), dtype=torch.float32, device=a2.device) if want_dmix else None
    with torch.cuda.device(a2.device):
        st = L.lib().bdp_head_fc3_bwd(L.ptr(dy), L.ptr(a2), a2.shape[1], L.ptr(w3), L.ptr(b3),
                                      L.ptr(mix), B, H, O, N2, L.ptr(da2), L.ptr(dw3), L.ptr(db3),
                                      L.ptr(dmix), L.stream_ptr())
    L.check(st, "bdp_head_fc3_bwd")
    return da2, dw3, db3, dmix


def sum_slabs(parts, n, S, stride, out):
    with torch.cuda.device(parts.device):
        st = L.lib().bdp_sum_slabs(L.ptr(parts), n, S, stride, L.ptr(out), L.stream_ptr())
    L.check(st, "bdp_sum_slabs")
    return out


# ------------------------------------------------------------------------------------------------
# the stacked 3-layer head: forward + backward
# ------------------------------------------------------------------------------------------------
class _HeadFn(torch.autograd.Function):
    """y = heads(x; stacked params, mix).  Inputs: x [B,N0], mix [B,Hm] (Hm heads per output group),
    then the stacked parameters.  Two output groups share fc1/fc2 machinery when `split` is given:
    heads [0, split) produce y1 with fc3 `w3a`, heads [split, H) produce y2 with `w3b` (the bin and
    res model lists of OneBinDeltaModel, which see the same features)."""

    @staticmethod
    def forward(ctx, x, mix, w1, g1, be1, w2, g2, be2, w3a, b3a, w3b, b3b, rm1, rv1, rm2, rv2,
                training):
        H, N1, N0 = w1.shape
        N2 = w2.shape[1]
        B = x.shape[0]
        dev = x.device
        ldb = _pad4(B)
        x = x.contiguous()
        if x.shape[1] % 4 != 0:
            raise ValueError("head: feature width must be a multiple of 4 (got %d)" % x.shape[1])
        # fc1: H1^T [H*N1, ldb] = W1 [H*N1, N0] (K-major) x X [B, N0] (K-major)
        h1 = torch.empty((H * N1, ldb), dtype=torch.float32, device=dev)
        gemm_tf32(w1, 0, N0, 0, x, 0, N0, 0, h1, 0, ldb, 0, H * N1, B, N0)
        a1, m1, is1 = bn_relu_fwd(h1, B, g1.reshape(-1), be1.reshape(-1),
                                  None if rm1 is None else rm1.view(-1),
                                  None if rv1 is None else rv1.view(-1), training)
        # fc2 (grouped): H2^T_g [N2, ldb] = W2_g [N2, N1] (K-major) x A1^T_g [N1, ldb] (MN-major)
        h2 = torch.empty((H * N2, ldb), dtype=torch.float32, device=dev)
        gemm_tf32(w2, 0, N1, N2 * N1, a1, 1, ldb, N1 * ldb, h2, 0, ldb, N2 * ldb, N2, B, N1, G=H)
        a2, m2, is2 = bn_relu_fwd(h2, B, g2.reshape(-1), be2.reshape(-1),
                                  None if rm2 is None else rm2.view(-1),
                                  None if rv2 is None else rv2.view(-1), training)
        Ha = w3a.shape[0]
        mix = mix.contiguous().float()
        y1 = fc3_fwd(a2[:Ha * N2], w3a, b3a, mix, B)
        y2 = fc3_fwd(a2[Ha * N2:], w3b, b3b, mix, B) if w3b is not None else None
        ctx.save_for_backward(x, mix, w1, g1, w2, g2, w3a, b3a, w3b, b3b, h1, a1, m1, is1, h2, a2, m2,
                              is2, rm1, rv1, rm2, rv2)
        ctx.training = training
        ctx.dims = (H, N0, N1, N2, B, ldb, Ha)
        if y2 is None:
            return y1
        return y1, y2

    @staticmethod
    def backward(ctx, dy1, dy2=None):
        (x, mix, w1, g1, w2, g2, w3a, b3a, w3b, b3b, h1, a1, m1, is1, h2, a2, m2, is2, rm1, rv1, rm2,
         rv2) = ctx.saved_tensors
        H, N0, N1, N2, B, ldb, Ha = ctx.dims
        training = ctx.training
        dev = x.device
        want_dmix = ctx.needs_input_grad[1]
        if not training:
            m1, is1 = rm1.view(-1), torch.rsqrt(rv1.view(-1) + BN_EPS)
            m2, is2 = rm2.view(-1), torch.rsqrt(rv2.view(-1) + BN_EPS)
        # fc3 backward
        da2 = torch.empty_like(a2)
        dy1 = dy1.contiguous().float()
        da2a, dw3a, db3a, dmix = fc3_bwd(dy1, a2[:Ha * N2], w3a, b3a, mix, B, want_dmix)
        da2[:Ha * N2] = da2a
        dw3b = db3b = None
        if w3b is not None:
            dy2 = dy2.contiguous().float()
            da2b, dw3b, db3b, dmix_b = fc3_bwd(dy2, a2[Ha * N2:], w3b, b3b, mix, B, want_dmix)
            da2[Ha * N2:] = da2b
            if want_dmix:
                dmix = dmix + dmix_b
        # bn2 backward
        dh2, dg2, dbe2 = bn_relu_bwd(da2, a2, h2, g2.reshape(-1), m2, is2, B, training)
        # fc2 wgrad: dW2_g [N2, N1] = dH2^T_g [N2, B] (K-major over the batch) x A1^T_g [N1, B] (K-major)
        dw2 = torch.empty_like(w2)
        gemm_tf32(dh2, 0, ldb, N2 * ldb, a1, 0, ldb, N1 * ldb, dw2, 0, N1, N2 * N1, N2, N1, B, G=H)
        # fc2 dgrad: dA1^T_g [N1, ldb] = W2_g^T (MN-major: [N2 rows, N1 contiguous]) x dH2^T_g (MN-major)
        da1 = torch.empty_like(a1)
        gemm_tf32(w2, 1, N1, N2 * N1, dh2, 1, ldb, N2 * ldb, da1, 0, ldb, N1 * ldb, N1, B, N2, G=H)
        # bn1 backward
        dh1, dg1, dbe1 = bn_relu_bwd(da1, a1, h1, g1.reshape(-1), m1, is1, B, training)
        # fc1 wgrad: dW1 [H*N1, N0] = dH1^T [H*N1, B] (K-major) x X [B, N0] (MN-major: batch rows)
        dw1 = torch.empty_like(w1)
        gemm_tf32(dh1, 0, ldb, 0, x, 1, N0, 0, dw1, 0, N0, 0, H * N1, N0, B)
        dx = None
        if ctx.needs_input_grad[0]:
            # fc1 dgrad: dX [B, N0]: D[m = feature, n = sample] = sum_k W1[k, m] dH1^T[k, n], split-K
            KK = H * N1
            splits = gemm_splits(KK, max(1, min(32, (L.lib().bdp_sm_count() * 128) // max(N0, 1))))
            parts = torch.empty((splits, B, N0), dtype=torch.float32, device=dev)
            gemm_tf32(w1, 1, N0, 0, dh1, 1, ldb, 0, parts, 1, N0, 0, N0, B, KK, splits=splits,
                      c_ss=B * N0)
            dx = torch.empty((B, N0), dtype=torch.float32, device=dev)
            sum_slabs(parts, B * N0, splits, B * N0, dx)
        return (dx, dmix, dw1, dg1.view_as(g1), dbe1.view_as(g1), dw2, dg2.view_as(g2),
                dbe2.view_as(g2), dw3a, db3a, dw3b, db3b, None, None, None, None, None)


def head_forward(x, mix, params, training):
    """params: dict with stacked tensors w1,g1,be1,w2,g2,be2,w3a,b3a,(w3b,b3b),rm1,rv1,rm2,rv2."""
    return _HeadFn.apply(x, mix, params["w1"], params["g1"], params["be1"], params["w2"],
                         params["g2"], params["be2"], params["w3a"], params["b3a"],
                         params.get("w3b"), params.get("b3b"), params["rm1"], params["rv1"],
                         params["rm2"], params["rv2"], training)
