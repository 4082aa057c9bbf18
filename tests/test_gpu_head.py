"""The bin-delta head kernels (tcgen05 TF32 GEMM, BatchNorm, fc3 mixing) against torch / the oracle.
Tolerance for anything that passes through the TF32 tensor-core GEMM: 2e-3 relative to the tensor's
scale (north_star: "2e-3 relative (bf16 head GEMM)"; TF32 keeps 3 more mantissa bits than bf16).
CUDA-core pieces (BatchNorm, fc3): 1e-5."""
import numpy as np
import pytest
import torch

import bdpose_oracle as O

pytestmark = pytest.mark.gpu
TF32_TOL = 2e-3


def scale_close(a, b, tol, msg=""):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    err = float((a - b).abs().max())
    ref = float(b.abs().max()) + 1e-30
    assert err <= tol * ref, "%s: max err %.3e vs scale %.3e (rel %.2e > %.1e)" % (msg, err, ref, err / ref, tol)


def _operand(rows, K, G, major, dev, gen):
    """random operand; returns (tensor as stored, logical [G, rows, K] view)"""
    if major == 0:
        t = torch.randn(G, rows, K, device=dev, generator=gen)
        return t, t
    t = torch.randn(G, K, rows, device=dev, generator=gen)
    return t, t.transpose(1, 2)


@pytest.mark.parametrize("M,N,K,G,am,bm,cl,splits", [
    (128, 32, 64, 1, 0, 0, 0, 1),
    (256, 32, 2048, 1, 0, 0, 0, 1),          # fc1 shape class
    (500, 32, 1000, 3, 0, 1, 0, 1),          # fc2 (grouped, B operand MN-major, ragged M and K)
    (500, 1000, 32, 2, 0, 0, 0, 1),          # fc2 wgrad (K = batch)
    (1000, 32, 500, 2, 1, 1, 0, 1),          # fc2 dgrad (both MN-major)
    (384, 2048, 48, 1, 0, 1, 0, 1),          # fc1 wgrad (B operand = X, MN-major), N tiles of 256
    (2048, 48, 3000, 1, 1, 1, 1, 5),         # fc1 dgrad: split-K partials, transposed store
    (200, 96, 100, 2, 0, 0, 1, 1),
    (132, 260, 36, 1, 1, 0, 0, 1),
])
def test_gemm_tf32(cuda, M, N, K, G, am, bm, cl, splits):
    from bdpose import head
    gen = torch.Generator(device=cuda).manual_seed(M + N + K)
    A, Al = _operand(M, K, G, am, cuda, gen)
    B, Bl = _operand(N, K, G, bm, cuda, gen)
    ref = torch.einsum("gmk,gnk->gmn", Al.double(), Bl.double())
    S = head.gemm_splits(K, splits)
    shape = (S, G, M, N) if cl == 0 else (S, G, N, M)
    C = torch.full(shape, float("nan"), device=cuda)
    a_ld = K if am == 0 else M
    b_ld = K if bm == 0 else N
    head.gemm_tf32(A, am, a_ld, M * K, B, bm, b_ld, N * K, C, cl, N if cl == 0 else M, M * N, M, N, K,
                   G=G, splits=splits, c_ss=G * M * N)
    out = C.sum(0)
    if cl == 1:
        out = out.transpose(1, 2)
    assert not torch.isnan(out).any()
    # TF32: error ~ 2^-11 * sqrt(K) * |a||b|; compare against the result scale
    scale_close(out, ref, TF32_TOL, "gemm")


def test_gemm_shared_operand_and_padding(cuda):
    """B operand shared by all groups (gstride 0) and a padded leading dimension on C."""
    from bdpose import head
    gen = torch.Generator(device=cuda).manual_seed(3)
    G, M, N, K, ldc = 3, 200, 40, 96, 44
    A = torch.randn(G, M, K, device=cuda, generator=gen)
    B = torch.randn(N, K, device=cuda, generator=gen)
    C = torch.zeros(G, M, ldc, device=cuda)
    head.gemm_tf32(A, 0, K, M * K, B, 0, K, 0, C, 0, ldc, M * ldc, M, N, K, G=G)
    ref = torch.einsum("gmk,nk->gmn", A.double(), B.double())
    scale_close(C[:, :, :N], ref, TF32_TOL, "shared-B")
    assert float(C[:, :, N:].abs().max()) == 0.0


def test_bn_relu_vs_torch(cuda):
    from bdpose import head
    torch.manual_seed(0)
    F, B, ldb = 300, 37, 40
    h = torch.randn(B, F, device=cuda, dtype=torch.float32) * 2 + 0.5
    bn = torch.nn.BatchNorm1d(F).to(cuda)
    bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.uniform_(-0.5, 0.5)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    hx = h.clone().requires_grad_(True)
    y = torch.relu(bn(hx))
    w = torch.randn_like(y)
    (y * w).sum().backward()
    hT = torch.zeros(F, ldb, device=cuda); hT[:, :B] = h.t()
    rm, rv = rm0.clone(), rv0.clone()
    a, mean, invstd = head.bn_relu_fwd(hT, B, bn.weight.data, bn.bias.data, rm, rv, True)
    torch.testing.assert_close(a[:, :B].t(), y.detach(), rtol=1e-5, atol=1e-6)
    assert float(a[:, B:].abs().max()) == 0
    torch.testing.assert_close(rm, bn.running_mean, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(rv, bn.running_var, rtol=1e-6, atol=1e-7)
    daT = torch.zeros(F, ldb, device=cuda); daT[:, :B] = w.t()
    dh, dg, db = head.bn_relu_bwd(daT, a, hT, bn.weight.data, mean, invstd, B, True)
    torch.testing.assert_close(dh[:, :B].t(), hx.grad, rtol=1e-4, atol=1e-5 * float(hx.grad.abs().max()))
    torch.testing.assert_close(dg, bn.weight.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(db, bn.bias.grad, rtol=1e-4, atol=1e-5)
    # eval mode
    bn.eval()
    ye = torch.relu(bn(h))
    ae, _, _ = head.bn_relu_fwd(hT, B, bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, False)
    torch.testing.assert_close(ae[:, :B].t(), ye, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("soft", [False, True])
def test_fc3_mix_vs_torch(cuda, soft):
    from bdpose import head
    torch.manual_seed(1)
    H, O, N2, B, ldb = 5, 37, 52, 9, 12
    a2 = torch.rand(H, N2, B, device=cuda)
    w3 = (torch.randn(H, O, N2, device=cuda) * 0.2).requires_grad_(True)
    b3 = torch.randn(H, O, device=cuda).requires_grad_(True)
    if soft:
        mix = torch.softmax(torch.randn(B, H, device=cuda), 1)
    else:
        mix = torch.zeros(B, H, device=cuda).scatter_(1, torch.randint(0, H, (B, 1), device=cuda), 1.0)
    mix.requires_grad_(True)
    a2r = a2.clone().requires_grad_(True)
    yh = torch.einsum("hoj,hjb->bho", w3, a2r) + b3[None]
    y = (yh * mix[:, :, None]).sum(1)
    dy = torch.randn_like(y)
    (y * dy).sum().backward()
    a2p = torch.zeros(H * N2, ldb, device=cuda); a2p[:, :B] = a2.reshape(H * N2, B)
    yg = head.fc3_fwd(a2p, w3.detach(), b3.detach(), mix.detach(), B)
    torch.testing.assert_close(yg, y.detach(), rtol=1e-5, atol=1e-5)
    da2, dw3, db3, dmix = head.fc3_bwd(dy, a2p, w3.detach(), b3.detach(), mix.detach(), B, True)
    torch.testing.assert_close(da2[:, :B].reshape(H, N2, B), a2r.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(dw3, w3.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(db3, b3.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(dmix, mix.grad, rtol=1e-5, atol=1e-5)
