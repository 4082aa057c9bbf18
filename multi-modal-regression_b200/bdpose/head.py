"""The category-conditioned bin-delta head on the B200 kernels.

`HeadStack` owns the STACKED parameters of H identical 3-layer MLPs (fc1 [H,N1,N0], bn1 [H,N1],
fc2 [H,N2,N1], bn2 [H,N2], fc3 [H,O,N2] + bias) and runs all of them on one feature batch:

  fc1   one tcgen05 TF32 GEMM over the flattened [H*N1, N0] weights        (bdp_gemm_tf32)
  bn1   batch statistics + affine + ReLU on batch-major activations         (bdp_bn_relu_fwd)
  fc2   grouped tcgen05 GEMM, one [N2,N1] weight per head                   (bdp_gemm_tf32, G=H)
  bn2   as bn1
  fc3   label-selected / soft-mixed output layer                            (bdp_head_fc3_fwd)

and the hand-derived backward (dgrad / wgrad through the same GEMM kernel with MN-major operand
descriptors, BatchNorm backward, fc3 backward).  Every head sees every sample in train mode, exactly
as the reference does (binDeltaModels.py:114-115): BatchNorm couples the batch inside each head, so
routing samples to their own category's head only would change the statistics.

Precision: the GEMMs run on the tensor cores straight from the fp32 master weights.  The default
"fp32" mode accumulates three TF32 MMAs per k-step on hi/lo operand splits (fp32-class results,
tolerance 1e-5 of the tensor scale in the tests); "tf32" issues one MMA per k-step — the
reduced-precision head GEMM of the north star (tolerance 2e-3), tighter than bf16.
"""
import ctypes as C
import os

import torch

from . import _lib as L

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# Head GEMM precision: "fp32" = 3xTF32 split accumulation (parity mode, default), "tf32" = one TF32
# MMA per k-step (the reduced-precision head GEMM; ~1e-3 relative, a few ReLU masks may flip).
PRECISION = os.environ.get("BDPOSE_HEAD_PRECISION", "fp32")


def set_precision(mode):
    global PRECISION
    if mode not in ("fp32", "tf32"):
        raise NameError("Unknown head precision passed")
    PRECISION = mode


def gemm_tf32(A, a_major, a_ld, a_gs, B, b_major, b_ld, b_gs, C, c_layout, ldc, c_gs, M, N, K, G=1,
              splits=1, c_ss=0, precise=None):
    """Raw launch of bdp_gemm_tf32 (see include/bdpose.h for the operand conventions)."""
    if precise is None:
        precise = PRECISION == "fp32"
    with torch.cuda.device(A.device):
        st = L.lib().bdp_gemm_tf32(L.ptr(A), a_major, a_ld, a_gs, L.ptr(B), b_major, b_ld, b_gs,
                                   L.ptr(C), c_layout, ldc, c_gs, M, N, K, G, splits, c_ss,
                                   1 if precise else 0, L.stream_ptr())
    L.check(st, "bdp_gemm_tf32")


def gemm_splits(K, splits):
    return L.lib().bdp_gemm_tf32_splits(K, splits)


def bn_relu_fwd(h, gamma, beta, running_mean, running_var, training, eps=BN_EPS,
                momentum=BN_MOMENTUM):
    """h [B, F] batch-major -> (a [B, F], save_mean [F], save_invstd [F])."""
    B, F = h.shape
    a = torch.empty_strided(h.shape, h.stride(), dtype=h.dtype, device=h.device)   # same row pitch
    if training:
        mean = torch.empty(F, dtype=torch.float32, device=h.device)
        invstd = torch.empty(F, dtype=torch.float32, device=h.device)
    else:
        mean = invstd = None
    with torch.cuda.device(h.device):
        st = L.lib().bdp_bn_relu_fwd(L.ptr(h), F, B, h.stride(0), L.ptr(gamma), L.ptr(beta),
                                     L.ptr(running_mean), L.ptr(running_var), L.ptr(mean),
                                     L.ptr(invstd), eps, momentum, 1 if training else 0, L.ptr(a),
                                     L.stream_ptr())
    L.check(st, "bdp_bn_relu_fwd")
    return a, mean, invstd


def bn_relu_bwd(da, a, h, gamma, mean, invstd, training=True):
    B, F = h.shape
    if da.stride() != h.stride() or a.stride() != h.stride():
        raise RuntimeError("bn_relu_bwd: da, a and h must share the row pitch")
    dh = torch.empty_strided(h.shape, h.stride(), dtype=h.dtype, device=h.device)
    dgamma = torch.empty(F, dtype=torch.float32, device=h.device)
    dbeta = torch.empty(F, dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        st = L.lib().bdp_bn_relu_bwd(L.ptr(da), L.ptr(a), L.ptr(h), L.ptr(gamma), L.ptr(mean),
                                     L.ptr(invstd), F, B, h.stride(0), 1 if training else 0,
                                     L.ptr(dh), L.ptr(dgamma), L.ptr(dbeta), L.stream_ptr())
    L.check(st, "bdp_bn_relu_bwd")
    return dh, dgamma, dbeta


def fc3_fwd(a2, w3, b3, mix):
    """a2 [B, H*N2] (a column slice of the stacked activations: row pitch a2.stride(0))."""
    H, O, N2 = w3.shape
    B = a2.shape[0]
    y = torch.empty((B, O), dtype=torch.float32, device=a2.device)
    with torch.cuda.device(a2.device):
        st = L.lib().bdp_head_fc3_fwd(L.ptr(a2), a2.stride(0), L.ptr(w3), L.ptr(b3), L.ptr(mix), B,
                                      H, O, N2, L.ptr(y), L.stream_ptr())
    L.check(st, "bdp_head_fc3_fwd")
    return y


def fc3_bwd(dy, a2, w3, b3, mix, want_dmix, da2=None):
    """da2 (optional) is the column slice of the stacked gradient buffer to fill (same pitch as a2)."""
    H, O, N2 = w3.shape
    B = a2.shape[0]
    if da2 is None:
        da2 = torch.empty_like(a2)
    if da2.stride(0) != a2.stride(0):
        raise RuntimeError("fc3_bwd: da2 and a2 must share the row pitch")
    dw3 = torch.empty_like(w3)
    db3 = torch.empty_like(b3)
    dmix = torch.empty((B, H), dtype=torch.float32, device=a2.device) if want_dmix else None
    with torch.cuda.device(a2.device):
        st = L.lib().bdp_head_fc3_bwd(L.ptr(dy), L.ptr(a2), a2.stride(0), L.ptr(w3), L.ptr(b3),
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
# stacked parameter storage
# ------------------------------------------------------------------------------------------------
_SLOTS = (("w1", "fc1", "weight"), ("g1", "bn1", "weight"), ("be1", "bn1", "bias"),
          ("w2", "fc2", "weight"), ("g2", "bn2", "weight"), ("be2", "bn2", "bias"))
_BUFS = (("rm1", "bn1", "running_mean"), ("rv1", "bn1", "running_var"),
         ("rm2", "bn2", "running_mean"), ("rv2", "bn2", "running_var"),
         ("nb1", "bn1", "num_batches_tracked"), ("nb2", "bn2", "num_batches_tracked"))


def _no_stack():
    return None


def _sentinel(stack):
    """Version counters of the first and last head's fc1 weight.  The autograd Functions below read
    the CURRENT stacked weights in backward; an optimizer step (or load_state_dict) between forward
    and backward bumps every Parameter's counter, so two of the 336 are enough to refuse what stock
    torch refuses — checking all of them would cost more than the step's whole host budget."""
    hs = stack.heads
    return (hs[0]._modules["fc1"]._parameters["weight"]._version,
            hs[-1]._modules["fc1"]._parameters["weight"]._version)


def _check_sentinel(ctx):
    if _sentinel(ctx.stack) != ctx.sentinel:
        raise RuntimeError("one of the variables needed for gradient computation has been modified by an "
                           "inplace operation: the head weights changed between forward and backward "
                           "(optimizer.step() / load_state_dict before backward())")


class HeadStack:
    """Keeps the parameters of a list of identical 3-layer MLP modules (fc1, bn1, fc2, bn2, fc3) in
    STACKED device buffers that the grouped kernels consume, while every module keeps its own
    nn.Parameter objects (state_dict keys, optimizers and per-index calls are untouched): each
    Parameter's .data is a view into the stacked buffer.  The views are re-established lazily
    (`ensure`) whenever something replaced the storage (.cuda(), .to(), deepcopy ...).

    `groups` is a list of module lists; all modules share the fc1/fc2 shapes, modules of one group
    share the fc3 shape (OneBinDeltaModel: [bin_models, res_models])."""

    def __init__(self, groups):
        self.groups = [list(g) for g in groups]
        self.heads = [m for g in self.groups for m in g]
        self.buf = None
        self.grad = None
        self.anchor = None
        self.plists = None
        self.descs = {}
        self.ws = None
        self._expected = None
        self.gflat = None        # the flat buffer behind the stacked gradient tensors
        self.gviews = None       # key -> tuple of per-module views into the stacked gradient buffers
        self.stacked = None      # key -> nn.Parameter over the stacked weights (opt-in fast path)

    # A stack is a cache over its modules (device buffers, ctypes descriptors): copies and pickles
    # of a model drop it and rebuild lazily on the next forward.
    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (_no_stack, ())

    # -- storage ---------------------------------------------------------------------------------
    def _probe(self):
        """data_ptr of three tensors per head, read through the module dicts (nn.Module.__getattr__
        costs ~1 us per hop; this runs on every forward)."""
        out = []
        for m in self.heads:
            mods = m._modules
            out.append(mods["fc1"]._parameters["weight"].data_ptr())
            out.append(mods["bn1"]._buffers["running_mean"].data_ptr())
            out.append(mods["fc3"]._parameters["weight"].data_ptr())
        return out

    def _is_current(self):
        return self.buf is not None and self._probe() == self._expected

    def ensure(self):
        if self._is_current():
            return self.buf
        heads = self.heads
        dev = heads[0].fc1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("bdpose heads run on CUDA only (parameters are on %s); call .cuda()" % dev)
        if len(self.groups) > L.HEAD_MAX_GROUPS or len({len(g) for g in self.groups}) != 1:
            raise RuntimeError("head stack: 1..%d fc3 groups with the same number of heads each" % L.HEAD_MAX_GROUPS)
        for m in heads:
            for bn in (m.bn1, m.bn2):
                if bn.eps != BN_EPS or bn.momentum != BN_MOMENTUM or not bn.affine or not bn.track_running_stats:
                    raise RuntimeError("bdpose heads implement nn.BatchNorm1d with the defaults the reference "
                                       "uses (eps=1e-5, momentum=0.1, affine, running stats); got eps=%r "
                                       "momentum=%r" % (bn.eps, bn.momentum))
        buf = {}
        with torch.no_grad():
            for key, sub, name in _SLOTS:
                src = [getattr(getattr(m, sub), name) for m in heads]
                st = torch.stack([p.detach().to(dev, torch.float32) for p in src]).contiguous()
                for i, p in enumerate(src):
                    p.data = st[i]
                buf[key] = st
            for key, sub, name in _BUFS:
                src = [getattr(getattr(m, sub), name) for m in heads]
                st = torch.stack([b.detach().to(dev) for b in src]).contiguous()
                for i, m in enumerate(heads):
                    getattr(m, sub)._buffers[name] = st[i]
                buf[key] = st
            buf["w3"], buf["b3"] = [], []
            for g in self.groups:
                w = torch.stack([m.fc3.weight.detach().to(dev, torch.float32) for m in g]).contiguous()
                b = torch.stack([m.fc3.bias.detach().to(dev, torch.float32) for m in g]).contiguous()
                for i, m in enumerate(g):
                    m.fc3.weight.data = w[i]
                    m.fc3.bias.data = b[i]
                buf["w3"].append(w)
                buf["b3"].append(b)
        self.buf = buf
        self.grad = None
        self.gflat = None
        self.gviews = None
        self.stacked = None
        self.plists = self._param_lists()
        self.descs = {}
        self.ws = None
        self._expected = self._probe()
        if self.anchor is None or self.anchor.device != dev:
            self.anchor = torch.zeros(1, device=dev, requires_grad=True)
        return buf

    def desc(self, training, precise):
        """struct bdp_head_desc for the current storage (cached per (training, precise))."""
        key = (bool(training), bool(precise))
        d = self.descs.get(key)
        if d is None:
            buf = self.buf
            H, N1, N0 = buf["w1"].shape
            d = L.HeadDesc()
            d.H, d.N0, d.N1, d.N2 = H, N0, N1, buf["w2"].shape[1]
            d.n_groups = len(self.groups)
            d.training, d.precise = int(key[0]), int(key[1])
            for g, (w3, b3) in enumerate(zip(buf["w3"], buf["b3"])):
                d.group_heads[g], d.group_out[g] = w3.shape[0], w3.shape[1]
                d.w3[g], d.b3[g] = w3.data_ptr(), b3.data_ptr()
            for f in ("w1", "g1", "be1", "w2", "g2", "be2", "rm1", "rv1", "rm2", "rv2"):
                setattr(d, f, buf[f].data_ptr())
            d.eps, d.momentum = BN_EPS, BN_MOMENTUM
            self.descs[key] = d
        return d

    def workspace(self, n_floats, dev):
        if self.ws is None or self.ws.numel() < n_floats or self.ws.device != dev:
            self.ws = torch.empty(n_floats, dtype=torch.float32, device=dev)
        return self.ws

    # -- gradients ---------------------------------------------------------------------------------
    def _param_lists(self):
        out = {key: [getattr(getattr(m, sub), name) for m in self.heads] for key, sub, name in _SLOTS}
        for gi, g in enumerate(self.groups):
            out["w3_%d" % gi] = [m.fc3.weight for m in g]
            out["b3_%d" % gi] = [m.fc3.bias for m in g]
        return out

    def _keys(self):
        keys = ["w1", "g1", "be1", "w2", "g2", "be2"]
        for g in range(len(self.groups)):
            keys += ["w3_%d" % g, "b3_%d" % g]
        return keys

    def _stacked_tensor(self, key):
        if key.startswith("w3_"):
            return self.buf["w3"][int(key[3:])]
        if key.startswith("b3_"):
            return self.buf["b3"][int(key[3:])]
        return self.buf[key]

    def grad_buffers(self):
        """The persistent stacked gradient buffers (one per stacked parameter tensor) and the
        per-module views into them, created once per storage generation."""
        if self.grad is None:
            # ONE flat allocation: the data-parallel all-reduce is a single collective over it
            shapes = {k: self._stacked_tensor(k).shape for k in self._keys()}
            sizes = {k: (int(torch.Size(v).numel()) + 3) // 4 * 4 for k, v in shapes.items()}
            self.gflat = torch.zeros(sum(sizes.values()), dtype=torch.float32,
                                     device=self.buf["w1"].device)
            self.grad, off = {}, 0
            for k in self._keys():
                n = int(torch.Size(shapes[k]).numel())
                self.grad[k] = self.gflat[off:off + n].view(shapes[k])
                off += sizes[k]
            self.gviews = {k: g.unbind(0) for k, g in self.grad.items()}
        return self.grad

    def stacked_parameters(self):
        """Opt-in fast path for new training code: ONE nn.Parameter per stacked tensor (10 for
        OneBinDeltaModel instead of 336 per-module ones) sharing memory with the per-module
        Parameters.  After this call backward() leaves its gradients on these (and only these), so an
        optimizer built over them steps the same weights with 30x fewer tensors to visit."""
        self.ensure()
        if self.stacked is None:
            self.stacked = {k: torch.nn.Parameter(self._stacked_tensor(k)) for k in self._keys()}
        return [self.stacked[k] for k in self._keys()]

    def grads_are_fresh(self):
        """True when no gradient is currently held (first backward after zero_grad(set_to_none))."""
        if self.stacked is not None:
            return self.stacked["w1"].grad is None
        return self.plists["w1"][0].grad is None and self.plists["w1"][-1].grad is None

    def grads_are_mine(self):
        if self.grad is None:
            return False
        if self.stacked is not None:
            g = self.stacked["w1"].grad
            return g is not None and g.data_ptr() == self.grad["w1"].data_ptr()
        first, last = self.plists["w1"][0].grad, self.plists["w1"][-1].grad
        return first is not None and last is not None and \
            first.data_ptr() == self.grad["w1"].data_ptr() and \
            last.data_ptr() == self.grad["w1"][-1].data_ptr()

    def publish(self):
        """Point every Parameter's .grad at its view of the persistent stacked gradient buffers
        (which backward() has just filled).  Every head gets a dense gradient, zeros included, as in
        the reference.  NOTE: the buffers are reused by the next fresh backward, so a
        reference to an old `.grad` kept across zero_grad(set_to_none=True) sees the new values."""
        if self.stacked is not None:
            for k, p in self.stacked.items():
                p.grad = self.grad[k]
            return
        for key, plist in self.plists.items():
            for p, v in zip(plist, self.gviews[key]):
                p.grad = v

    def accumulate(self, grads):
        """Second backward before the next zero_grad (two forwards, one backward:
        learnGeodesicBDModel.py:116-120, 183), or foreign gradients already present."""
        if self.grads_are_mine():
            for key in self.grad:
                self.grad[key].add_(grads[key])
            return
        if self.stacked is not None:
            for k, p in self.stacked.items():
                if p.grad is None:
                    p.grad = grads[k]
                else:
                    p.grad.add_(grads[k])
            return
        for key, plist in self.plists.items():
            for p, v in zip(plist, grads[key].unbind(0)):
                if p.grad is None:
                    p.grad = v.clone()
                else:
                    p.grad.add_(v)


# ------------------------------------------------------------------------------------------------
# the stacked 3-layer head: forward + backward
# ------------------------------------------------------------------------------------------------
class _HeadFn(torch.autograd.Function):
    """(y_0, y_1, ...) = heads(x, mix) for a HeadStack: one output per fc3 group.  Differentiable
    inputs are x, mix and the stack's anchor (which only keeps the node alive when neither x nor mix
    needs a gradient); parameter gradients are deposited on the modules' Parameters directly."""

    @staticmethod
    def forward(ctx, x, mix, anchor, stack, training):
        buf = stack.ensure()
        H, N1, N0 = buf["w1"].shape
        B = x.shape[0]
        dev = x.device
        if x.dim() != 2 or x.shape[1] != N0:
            raise RuntimeError("head: input has %s features, fc1 expects %d" % (tuple(x.shape[1:]), N0))
        if training and B < 2:
            raise ValueError("Expected more than 1 value per channel when training, got input size "
                             "[%d, %d]" % (B, N1))
        Hg = buf["w3"][0].shape[0]
        if mix.dim() != 2 or mix.shape[0] != B or mix.shape[1] != Hg:
            raise RuntimeError("head: mixing weights are %s for a batch of %d and %d heads per group"
                               % (tuple(mix.shape), B, Hg))
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        mix = mix.detach()
        if mix.dtype != torch.float32 or not mix.is_contiguous():
            mix = mix.float().contiguous()
        desc = stack.desc(training, PRECISION == "fp32")
        lib = L.lib()
        dref = C.byref(desc)
        saved = torch.empty(lib.bdp_head_saved_floats(dref, B), dtype=torch.float32, device=dev)
        ys = [torch.empty((B, w3.shape[1]), dtype=torch.float32, device=dev) for w3 in buf["w3"]]
        yptr = (C.c_void_p * len(ys))(*[y.data_ptr() for y in ys])
        with torch.cuda.device(dev):
            st = lib.bdp_head_forward(dref, x.data_ptr(), mix.data_ptr(), B, saved.data_ptr(), yptr,
                                      L.stream_ptr())
        L.check(st, "bdp_head_forward")
        if training:
            buf["nb1"] += 1
            buf["nb2"] += 1
        ctx.stack = stack
        ctx.sentinel = _sentinel(stack)
        ctx.saved = (x, mix, saved)
        ctx.training = training
        ctx.precise = PRECISION == "fp32"
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        _check_sentinel(ctx)
        stack = ctx.stack
        buf = stack.buf
        x, mix, saved = ctx.saved
        B = x.shape[0]
        dev = x.device
        desc = stack.desc(ctx.training, ctx.precise)
        lib = L.lib()
        dref = C.byref(desc)
        dyl = []
        for dy, w3 in zip(dys, buf["w3"]):
            if dy is None:
                dy = torch.zeros((B, w3.shape[1]), dtype=torch.float32, device=dev)
            elif dy.dtype != torch.float32 or not dy.is_contiguous():
                dy = dy.float().contiguous()
            dyl.append(dy)
        n = len(dyl)
        ws = stack.workspace(lib.bdp_head_bwd_workspace_floats(dref, B), dev)
        # first backward since zero_grad: the kernels write the persistent stacked gradient buffers
        # directly; otherwise they write temporaries that are accumulated
        fresh = stack.grads_are_fresh()
        if fresh:
            grads = stack.grad_buffers()
        else:
            grads = {k: torch.empty_like(stack._stacked_tensor(k)) for k in stack._keys()}
        dmix = torch.empty_like(mix) if ctx.needs_input_grad[1] else None
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        arr = C.c_void_p * n
        with torch.cuda.device(dev):
            st = lib.bdp_head_backward(
                dref, x.data_ptr(), mix.data_ptr(), B, saved.data_ptr(),
                arr(*[t.data_ptr() for t in dyl]), ws.data_ptr(),
                grads["w1"].data_ptr(), grads["g1"].data_ptr(), grads["be1"].data_ptr(),
                grads["w2"].data_ptr(), grads["g2"].data_ptr(), grads["be2"].data_ptr(),
                arr(*[grads["w3_%d" % g].data_ptr() for g in range(n)]),
                arr(*[grads["b3_%d" % g].data_ptr() for g in range(n)]),
                L.ptr(dmix), L.ptr(dx), L.stream_ptr())
        L.check(st, "bdp_head_backward")
        if fresh:
            stack.publish()
        else:
            stack.accumulate(grads)
        return dx, dmix, None, None, None


# ------------------------------------------------------------------------------------------------
# dense outputs of every head (scripts that loop `bin_models[i](x)` themselves)
# ------------------------------------------------------------------------------------------------
class _HeadAllFn(torch.autograd.Function):
    """(Y_0 [B, Hg, O_0], Y_1 [B, Hg, O_1], ...) = the UNMIXED outputs of every head of a HeadStack
    on x.  fc1 / fc2 run as the stacked tcgen05 GEMMs, BatchNorm through bn_relu, the output layers as
    one batched matmul per fc3 group."""

    @staticmethod
    def forward(ctx, x, anchor, stack, training):
        buf = stack.ensure()
        w1, w2 = buf["w1"], buf["w2"]
        H, N1, N0 = w1.shape
        N2 = w2.shape[1]
        B = x.shape[0]
        F1, F2 = H * N1, H * N2
        if x.dim() != 2 or x.shape[1] != N0:
            raise RuntimeError("head: input has %s features, fc1 expects %d" % (tuple(x.shape[1:]), N0))
        if training and B < 2:
            raise ValueError("Expected more than 1 value per channel when training, got input size "
                             "[%d, %d]" % (B, N1))
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        dev = x.device
        h1 = torch.empty((B, F1), dtype=torch.float32, device=dev)
        gemm_tf32(x, 0, N0, 0, w1, 0, N0, 0, h1, 0, F1, 0, B, F1, N0)
        a1, m1, is1 = bn_relu_fwd(h1, buf["g1"].view(-1), buf["be1"].view(-1), buf["rm1"].view(-1),
                                  buf["rv1"].view(-1), training)
        h2 = torch.empty((B, F2), dtype=torch.float32, device=dev)
        gemm_tf32(a1, 0, F1, N1, w2, 0, N1, N2 * N1, h2, 0, F2, N2, B, N2, N1, G=H)
        a2, m2, is2 = bn_relu_fwd(h2, buf["g2"].view(-1), buf["be2"].view(-1), buf["rm2"].view(-1),
                                  buf["rv2"].view(-1), training)
        if training:
            buf["nb1"] += 1
            buf["nb2"] += 1
        else:
            m1, is1 = buf["rm1"].view(-1).clone(), torch.rsqrt(buf["rv1"].view(-1) + BN_EPS)
            m2, is2 = buf["rm2"].view(-1).clone(), torch.rsqrt(buf["rv2"].view(-1) + BN_EPS)
        a2h = a2.view(B, H, N2).transpose(0, 1)                          # [H, B, N2]
        ys, off = [], 0
        for w3, b3 in zip(buf["w3"], buf["b3"]):
            Hg = w3.shape[0]
            y = torch.baddbmm(b3.unsqueeze(1), a2h[off:off + Hg], w3.transpose(1, 2))   # [Hg, B, O]
            ys.append(y.transpose(0, 1).contiguous())
            off += Hg
        ctx.stack, ctx.training = stack, training
        ctx.sentinel = _sentinel(stack)
        ctx.saved = (x, h1, a1, m1, is1, h2, a2, m2, is2)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        _check_sentinel(ctx)
        stack = ctx.stack
        buf = stack.buf
        x, h1, a1, m1, is1, h2, a2, m2, is2 = ctx.saved
        w1, w2 = buf["w1"], buf["w2"]
        H, N1, N0 = w1.shape
        N2 = w2.shape[1]
        B = x.shape[0]
        F1, F2 = H * N1, H * N2
        dev = x.device
        training = ctx.training
        a2h = a2.view(B, H, N2).transpose(0, 1)
        grads = {}
        da2h = torch.empty((H, B, N2), dtype=torch.float32, device=dev)
        off = 0
        for gi, (w3, b3) in enumerate(zip(buf["w3"], buf["b3"])):
            Hg, O = w3.shape[0], w3.shape[1]
            dy = dys[gi]
            dy = torch.zeros((B, Hg, O), device=dev) if dy is None else dy.float()
            dyh = dy.transpose(0, 1).contiguous()                                  # [Hg, B, O]
            grads["w3_%d" % gi] = torch.bmm(dyh.transpose(1, 2), a2h[off:off + Hg])
            grads["b3_%d" % gi] = dyh.sum(1)
            torch.bmm(dyh, w3, out=da2h[off:off + Hg])
            off += Hg
        da2 = da2h.transpose(0, 1).reshape(B, F2).contiguous()
        dh2, dg2, dbe2 = bn_relu_bwd(da2, a2, h2, buf["g2"].view(-1), m2, is2, training)
        dw2 = torch.empty_like(w2)
        gemm_tf32(dh2, 1, F2, N2, a1, 1, F1, N1, dw2, 0, N1, N2 * N1, N2, N1, B, G=H)
        da1 = torch.empty_like(a1)
        gemm_tf32(dh2, 0, F2, N2, w2, 1, N1, N2 * N1, da1, 0, F1, N1, B, N1, N2, G=H)
        dh1, dg1, dbe1 = bn_relu_bwd(da1, a1, h1, buf["g1"].view(-1), m1, is1, training)
        dw1 = torch.empty_like(w1)
        gemm_tf32(dh1, 1, F1, 0, x, 1, N0, 0, dw1, 0, N0, 0, F1, N0, B)
        grads.update(w1=dw1, g1=dg1.view(H, N1), be1=dbe1.view(H, N1), w2=dw2, g2=dg2.view(H, N2),
                     be2=dbe2.view(H, N2))
        if stack.grads_are_fresh():
            gb = stack.grad_buffers()
            for k in gb:
                gb[k].copy_(grads[k])
            stack.publish()
        else:
            stack.accumulate(grads)
        dx = None
        if ctx.needs_input_grad[0]:
            n_tiles = (N0 + 255) // 256
            splits = gemm_splits(F1, max(1, min(64, L.lib().bdp_sm_count() // n_tiles)))
            parts = torch.empty((splits, B, N0), dtype=torch.float32, device=dev)
            gemm_tf32(dh1, 0, F1, 0, w1, 1, N0, 0, parts, 0, N0, 0, B, N0, F1, splits=splits, c_ss=B * N0)
            dx = torch.empty((B, N0), dtype=torch.float32, device=dev)
            sum_slabs(parts, B * N0, splits, B * N0, dx)
        return dx, None, None, None


def run_heads_all(stack, x, training):
    """Unmixed outputs of every head: one [B, Hg, O_g] tensor per fc3 group."""
    stack.ensure()
    return _HeadAllFn.apply(x, stack.anchor, stack, training)


MEMO = os.environ.get("BDPOSE_HEAD_MEMO", "1") != "0"


class HeadFamily:
    """The sibling heads of one model (OneBinDeltaModel: bin_models + res_models).  Scripts that define
    their own forward call `self.bin_models[i](x)` head by head with the same x
    (learnJointCatPoseModel_weighted.py:112-113, evaluateJointModel.py:86-96): the first such call
    runs ALL heads fused (`run_heads_all`) and the siblings' calls return their slices, which is the
    same work the reference does in its loop — every head sees every sample — in 1/24 of the launches.
    One cached run serves one round: each head at most once, same input tensor (object, version,
    storage) and mode, that head's parameters untouched since the run.  BDPOSE_HEAD_MEMO=0 turns this off (every call then runs as a one-head
    stack)."""

    def __init__(self, lists, two_layer=False):
        self.lists = lists                 # the nn.ModuleLists, in fc3-group order
        self.two_layer = two_layer         # res_2layer siblings (Mlp2Stack) instead of 3-layer heads
        self._stack = None
        self._key = None
        self._x_ref = None
        self._outs = None
        self._pver = {}
        self._served = set()
        self._where = {}

    def __deepcopy__(self, memo):
        import copy
        new = HeadFamily.__new__(HeadFamily)
        memo[id(self)] = new        # the member modules point back at the family: one copy for all
        new.__init__([copy.deepcopy(l, memo) for l in self.lists], self.two_layer)
        return new

    def __reduce__(self):
        return (HeadFamily, (self.lists, self.two_layer))

    def stack(self):
        groups = [list(l) for l in self.lists]
        st = self._stack
        flat = [m for g in groups for m in g]
        if st is None or len(st.heads) != len(flat) or any(a is not b for a, b in zip(st.heads, flat)):
            st = Mlp2Stack(flat) if self.two_layer else HeadStack(groups)
            self._stack = st
            self._key = None
        return st

    @staticmethod
    def _versions(m):
        return tuple(p._version for sub in m._modules.values() for p in sub._parameters.values()
                     if p is not None)

    def output_of(self, module, x, training):
        key = (id(x), x._version, x.data_ptr(), tuple(x.shape), bool(training), torch.is_grad_enabled(),
               x.requires_grad)
        # One cached run serves ONE round of the script's loop: every head at most once, for the same
        # input, with its parameters untouched since the run.  A head asked a second time starts a
        # new round (as the reference would recompute — and, in train mode, move the BatchNorm
        # statistics — on every call).
        fresh = key != self._key or self._x_ref is not x or id(module) in self._served or \
            self._pver.get(id(module)) != self._versions(module)
        if fresh:
            st = self.stack()          # membership / storage checks: once per round, not per head
            if self.two_layer:
                self._outs = (run_mlp2_all(st, x, training),)
                self._where = {id(m): (0, j) for j, m in enumerate(st.heads)}
            else:
                self._outs = run_heads_all(st, x, training)
                self._where = {id(m): (gi, j) for gi, g in enumerate(st.groups) for j, m in enumerate(g)}
            self._key, self._x_ref = key, x
            self._pver = {id(m): self._versions(m) for m in st.heads}
            self._served = set()
        self._served.add(id(module))
        gi, j = self._where[id(module)]
        return self._outs[gi][:, j, :]


# ------------------------------------------------------------------------------------------------
# stacks of two-layer heads (one delta per bin: SURVEY §8(f)-1)
# ------------------------------------------------------------------------------------------------
class Mlp2Stack:
    """H identical two-layer heads fc2(relu(bn1(fc1 x))) (binDeltaModels.res_2layer, 49-59) over
    stacked storage: fc1 [H,N1,N0] streams through the tcgen05 GEMM (for OneDeltaPerBinModel C*K = 192
    heads: 157 MB of weights), BatchNorm through bn_relu, the tiny per-head output layer through
    a batched matmul.  Same caching rules as HeadStack (Parameters are views of the stacked buffers)."""
    _slots = (("w1", "fc1", "weight"), ("g1", "bn1", "weight"), ("be1", "bn1", "bias"),
              ("w2", "fc2", "weight"), ("b2", "fc2", "bias"))
    _bufs = (("rm1", "bn1", "running_mean"), ("rv1", "bn1", "running_var"),
             ("nb1", "bn1", "num_batches_tracked"))

    def __init__(self, heads):
        self.heads = list(heads)
        self.buf = None
        self.grad = None
        self.gviews = None
        self._expected = None

    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (_no_stack, ())

    def _probe(self):
        out = []
        for m in self.heads:
            mods = m._modules
            out.append(mods["fc1"]._parameters["weight"].data_ptr())
            out.append(mods["bn1"]._buffers["running_mean"].data_ptr())
            out.append(mods["fc2"]._parameters["weight"].data_ptr())
        return out

    def ensure(self):
        if self.buf is not None and self._probe() == self._expected:
            return self.buf
        dev = self.heads[0].fc1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("bdpose heads run on CUDA only (parameters are on %s); call .cuda()" % dev)
        buf = {}
        with torch.no_grad():
            for key, sub, name in self._slots:
                src = [getattr(getattr(m, sub), name) for m in self.heads]
                st = torch.stack([p.detach().to(dev, torch.float32) for p in src]).contiguous()
                for i, p in enumerate(src):
                    p.data = st[i]
                buf[key] = st
            for key, sub, name in self._bufs:
                src = [getattr(getattr(m, sub), name) for m in self.heads]
                st = torch.stack([b.detach().to(dev) for b in src]).contiguous()
                for i, m in enumerate(self.heads):
                    getattr(m, sub)._buffers[name] = st[i]
                buf[key] = st
        self.buf = buf
        self.grad = None
        self.gviews = None
        self.plists = {key: [getattr(getattr(m, sub), name) for m in self.heads]
                       for key, sub, name in self._slots}
        self._expected = self._probe()
        return buf

    def publish_or_accumulate(self, grads):
        """First backward since zero_grad: the Parameters' .grad become views of the persistent
        stacked buffers (filled with `grads`); otherwise accumulate."""
        first = self.plists["w1"][0].grad
        if self.grad is None:
            self.grad = {k: torch.empty_like(self.buf[k]) for k, _, _ in self._slots}
            self.gviews = {k: g.unbind(0) for k, g in self.grad.items()}
        mine = first is not None and first.data_ptr() == self.grad["w1"].data_ptr()
        if first is None:
            for k in self.grad:
                self.grad[k].copy_(grads[k])
            for k, plist in self.plists.items():
                for p, v in zip(plist, self.gviews[k]):
                    p.grad = v
        elif mine:
            for k in self.grad:
                self.grad[k].add_(grads[k])
        else:
            for k, plist in self.plists.items():
                for p, v in zip(plist, grads[k].unbind(0)):
                    p.grad = v.clone() if p.grad is None else p.grad.add_(v)


class _Mlp2Fn(torch.autograd.Function):
    """Y [B, H, O] = every head of an Mlp2Stack on x [B, N0]."""

    @staticmethod
    def forward(ctx, x, stack, training):
        buf = stack.ensure()
        H, N1, N0 = buf["w1"].shape
        O = buf["w2"].shape[1]
        B = x.shape[0]
        F1 = H * N1
        if x.dim() != 2 or x.shape[1] != N0:
            raise RuntimeError("head: input has %s features, fc1 expects %d" % (tuple(x.shape[1:]), N0))
        if N0 % 4 or N1 % 4:
            raise RuntimeError("head: layer widths must be multiples of 4 (got %d, %d)" % (N0, N1))
        if training and B < 2:
            raise ValueError("Expected more than 1 value per channel when training, got input size "
                             "[%d, %d]" % (B, N1))
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        h1 = torch.empty((B, F1), dtype=torch.float32, device=x.device)
        gemm_tf32(x, 0, N0, 0, buf["w1"], 0, N0, 0, h1, 0, F1, 0, B, F1, N0)
        a1, m1, is1 = bn_relu_fwd(h1, buf["g1"].view(-1), buf["be1"].view(-1), buf["rm1"].view(-1),
                                  buf["rv1"].view(-1), training)
        if not training:
            m1, is1 = buf["rm1"].view(-1).clone(), torch.rsqrt(buf["rv1"].view(-1) + BN_EPS)
        else:
            buf["nb1"] += 1
        # per-head output layer: [H, B, N1] x [H, N1, O] + bias
        y = torch.baddbmm(buf["b2"].unsqueeze(1), a1.view(B, H, N1).transpose(0, 1),
                          buf["w2"].transpose(1, 2))
        ctx.stack, ctx.training = stack, training
        ctx.sentinel = _sentinel(stack)
        ctx.saved = (x, h1, a1, m1, is1)
        return y.transpose(0, 1).contiguous()

    @staticmethod
    def backward(ctx, dy):
        _check_sentinel(ctx)
        stack = ctx.stack
        buf = stack.buf
        x, h1, a1, m1, is1 = ctx.saved
        H, N1, N0 = buf["w1"].shape
        B = x.shape[0]
        F1 = H * N1
        dyh = dy.float().transpose(0, 1).contiguous()                      # [H, B, O]
        a1h = a1.view(B, H, N1).transpose(0, 1)                            # [H, B, N1]
        grads = {"w2": torch.bmm(dyh.transpose(1, 2), a1h), "b2": dyh.sum(1)}
        da1 = torch.bmm(dyh, buf["w2"]).transpose(0, 1).reshape(B, F1).contiguous()
        dh1, dg1, dbe1 = bn_relu_bwd(da1, a1, h1, buf["g1"].view(-1), m1, is1, ctx.training)
        dw1 = torch.empty_like(buf["w1"])
        gemm_tf32(dh1, 1, F1, 0, x, 1, N0, 0, dw1, 0, N0, 0, F1, N0, B)
        grads.update(w1=dw1, g1=dg1.view(H, N1), be1=dbe1.view(H, N1))
        stack.publish_or_accumulate(grads)
        dx = None
        if ctx.needs_input_grad[0]:
            n_tiles = (N0 + 255) // 256
            splits = gemm_splits(F1, max(1, min(64, L.lib().bdp_sm_count() // n_tiles)))
            parts = torch.empty((splits, B, N0), dtype=torch.float32, device=x.device)
            gemm_tf32(dh1, 0, F1, 0, buf["w1"], 1, N0, 0, parts, 0, N0, 0, B, N0, F1, splits=splits,
                      c_ss=B * N0)
            dx = torch.empty((B, N0), dtype=torch.float32, device=x.device)
            sum_slabs(parts, B * N0, splits, B * N0, dx)
        return dx, None, None


def run_mlp2_all(stack, x, training):
    """Every head of an Mlp2Stack on features x [B, N0] -> [B, H, O]."""
    stack.ensure()
    if not x.requires_grad:
        x = x.detach().requires_grad_(True)      # keeps the node alive so parameter grads are produced
    return _Mlp2Fn.apply(x, stack, training)


def allreduce_stack_grads(stack, group=None, average=True):
    """Data-parallel step of the head: all-reduce the STACKED gradient buffers of a
    HeadStack in place — a handful of large contiguous tensors (fc1 alone is 197 MB for the Pascal
    head) instead of 336 per-module ones; the modules' `.grad` views see the reduced values.  Each
    rank ran its own batch (per-GPU BatchNorm statistics, DDP semantics)."""
    import torch.distributed as dist
    if stack.grad is None or not (dist.is_available() and dist.is_initialized()):
        return
    ws = dist.get_world_size(group)
    if ws == 1:
        return
    flat = getattr(stack, "gflat", None)
    if flat is not None and stack.grads_are_mine():
        # one collective over the flat gradient buffer; NCCL averages in the reduction itself
        if average and dist.get_backend(group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                flat.div_(ws)
        return
    # The Parameters' .grad are NOT views of the stacked buffers (foreign gradients were present and
    # accumulate() added into separate tensors, or the stack was re-published): reduce the gradients
    # the Parameters actually hold.
    plists = getattr(stack, "plists", None) or {}
    grads = [p.grad for plist in plists.values() for p in plist if p.grad is not None]
    if getattr(stack, "stacked", None):
        grads = [p.grad for p in stack.stacked.values() if p.grad is not None]
    if not grads:
        if plists:
            raise RuntimeError("allreduce_stack_grads: the stack's parameters hold no gradients")
        grads = [g for _, g in sorted(stack.grad.items())]     # a bare stack of gradient buffers
    works = [dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True) for g in grads]
    for w in works:
        w.wait()
    if average:
        for g in grads:
            g.div_(ws)


def sync_head_gradients(model, group=None):
    """All-reduce(mean) the head gradients of every fused stack found under `model` (the
    containers keep theirs in `_stack` / `_stack2`, sibling heads in their `_family`)."""
    seen = set()
    for m in model.modules():
        cands = [m.__dict__.get(name) for name in ("_stack", "_solo", "_stack2")]
        fam = m.__dict__.get("_family")
        if fam is not None:
            cands.append(fam._stack)
        for st in cands:
            if st is not None and id(st) not in seen and getattr(st, "grad", None) is not None:
                seen.add(id(st))
                allreduce_stack_grads(st, group)


def run_heads(stack, x, mix, training):
    """All heads of `stack` on features x [B, N0] with mixing weights mix [B, heads per group].
    Returns one [B, O_g] tensor per fc3 group."""
    stack.ensure()
    return _HeadFn.apply(x, mix, stack.anchor, stack, training)


def onehot(label, num_classes):
    """[B,1] (or [B]) int64 labels -> [B, C] float one-hot built on the device (the reference builds
    it on the CPU and copies it back every forward: binDeltaModels.py:116-117)."""
    label = label.reshape(-1, 1).long()
    return torch.zeros(label.shape[0], num_classes, device=label.device).scatter_(1, label, 1.0)
