"""The whole bin-delta head step as ONE CUDA graph (fixed batch size).

At the reference's batch sizes a training step of the heads is ~20 kernels of 5-50 us each: issued
one by one from Python the step is host-bound (0.45-0.57 ms for ~0.3 ms of GPU work).
`GraphedBinDeltaStep` captures

    heads forward  ->  fused bin-delta loss (forward + backward)  ->  heads backward

into a CUDA graph over static buffers: a step is four small input copies and one graph launch.  It is
an opt-in API for new training loops (the drop-in modules keep the eager autograd path); semantics
are those of `loss = Lc + w * Lr; loss.backward()` after `optimizer.zero_grad()`:

  * parameter gradients are written (not accumulated) into the stack's persistent stacked gradient
    buffers and published on the Parameters (per-module views, or the stacked Parameters when
    `model.stacked_head_parameters()` was requested);
  * `step.dx` holds d loss / d features for the trunk (`feat.backward(step.dx)`);
  * BatchNorm running statistics and `num_batches_tracked` advance once per call (training mode).

Reference: the step of learnGeodesicBDModel.py:156-205 (model forward 116-120, CE + geodesic loss
178-180, backward 183) on precomputed trunk features.
"""
import ctypes as C

import torch

from . import _lib as L
from . import head as H
from . import ops


class GraphedBinDeltaStep:
    def __init__(self, model, batch_size, keys, pose_mode=L.POSE_GEODESIC_AA, use_keys=True,
                 pose_weight=1.0, want_dx=True):
        stack = model._heads() if hasattr(model, "_heads") else model
        buf = stack.ensure()
        if len(stack.groups) != 2:
            raise RuntimeError("GraphedBinDeltaStep expects a bin group and a delta group of heads")
        self.stack = stack
        self.B = B = int(batch_size)
        dev = buf["w1"].device
        self.dev = dev
        Hh, N1, N0 = buf["w1"].shape
        Hg = buf["w3"][0].shape[0]
        K, nd = buf["w3"][0].shape[1], buf["w3"][1].shape[1]
        self.precise = H.PRECISION == "fp32"
        self.desc = stack.desc(True, self.precise)
        lib = L.lib()
        dref = C.byref(self.desc)
        f32 = dict(dtype=torch.float32, device=dev)
        tdim = 9 if pose_mode in (L.POSE_RIEMANNIAN, L.POSE_ROTMAT) else nd
        # static inputs
        self.x = torch.zeros((B, N0), **f32)
        self.label = torch.zeros((B, 1), dtype=torch.int64, device=dev)
        self.mix = torch.zeros((B, Hg), **f32)
        self.bins = torch.zeros(B, dtype=torch.int64, device=dev)
        self.target = torch.zeros((B, tdim), **f32)
        self.weight = torch.full((1,), float(pose_weight), **f32)
        self.keys = None if keys is None else keys.detach().to(dev, torch.float32).reshape(K, -1).contiguous()
        # static intermediates / outputs
        self.saved = torch.empty(lib.bdp_head_saved_floats(dref, B), **f32)
        self.y1 = torch.empty((B, K), **f32)
        self.y2 = torch.empty((B, nd), **f32)
        self.losses = torch.zeros(2, **f32)
        self.g_logits = torch.empty((B, K), **f32)
        self.g_pred = torch.empty((B, nd), **f32)
        self.dy2 = torch.empty((B, nd), **f32)
        self.argmax = torch.empty(B, dtype=torch.int64, device=dev)
        self.loss_ws = torch.zeros(lib.bdp_bd_loss_workspace_bytes(B), dtype=torch.uint8, device=dev)
        self.bwd_ws = torch.empty(lib.bdp_head_bwd_workspace_floats(dref, B), **f32)
        self.dx = torch.empty((B, N0), **f32) if want_dx else None
        self.grads = stack.grad_buffers()
        self.pose_mode, self.use_keys = int(pose_mode), 1 if use_keys else 0
        self._yptr = (C.c_void_p * 2)(self.y1.data_ptr(), self.y2.data_ptr())
        self._dyptr = (C.c_void_p * 2)(self.g_logits.data_ptr(), self.dy2.data_ptr())
        self._dw3 = (C.c_void_p * 2)(self.grads["w3_0"].data_ptr(), self.grads["w3_1"].data_ptr())
        self._db3 = (C.c_void_p * 2)(self.grads["b3_0"].data_ptr(), self.grads["b3_1"].data_ptr())
        self._storage_probe = stack._expected
        # warm-up on a side stream (function attributes, tensor-map cache, allocator), then capture
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        keep = {k: buf[k].clone() for k in ("rm1", "rv1", "rm2", "rv2", "nb1", "nb2")}
        with torch.cuda.stream(side):
            self._issue()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for k, v in keep.items():
            buf[k].copy_(v)                      # the warm-up must not move the running statistics
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._issue()

    def _issue(self):
        lib, st, buf = L.lib(), L.stream_ptr(), self.stack.buf
        dref = C.byref(self.desc)
        B = self.B
        self.mix.zero_()
        self.mix.scatter_(1, self.label, 1.0)
        with torch.cuda.device(self.dev):
            L.check(lib.bdp_head_forward(dref, self.x.data_ptr(), self.mix.data_ptr(), B,
                                         self.saved.data_ptr(), self._yptr, st), "bdp_head_forward")
            K, nd = self.y1.shape[1], self.y2.shape[1]
            L.check(lib.bdp_bd_loss_fwd_bwd(
                self.y1.data_ptr(), B, K, K, self.bins.data_ptr(), self.y2.data_ptr(), nd,
                L.ptr(self.keys), self.use_keys, self.target.data_ptr(), self.pose_mode,
                self.losses.data_ptr(), None, None, self.g_logits.data_ptr(), self.g_pred.data_ptr(),
                0.0, self.argmax.data_ptr(), self.loss_ws.data_ptr(), self.loss_ws.numel(), st),
                "bdp_bd_loss_fwd_bwd")
        torch.mul(self.g_pred, self.weight, out=self.dy2)          # d(Lc + w Lr)/d delta
        g = self.grads
        with torch.cuda.device(self.dev):
            L.check(lib.bdp_head_backward(
                dref, self.x.data_ptr(), self.mix.data_ptr(), B, self.saved.data_ptr(), self._dyptr,
                self.bwd_ws.data_ptr(), g["w1"].data_ptr(), g["g1"].data_ptr(), g["be1"].data_ptr(),
                g["w2"].data_ptr(), g["g2"].data_ptr(), g["be2"].data_ptr(), self._dw3, self._db3, None,
                L.ptr(self.dx), st), "bdp_head_backward")
        buf["nb1"] += 1
        buf["nb2"] += 1

    def __call__(self, feat, label, bins, target, pose_weight=None):
        """feat [B,N0], label [B,1] int64 class ids, bins [B] int64, target [B,ndim|9].
        Returns (Lc, Lr) as 0-dim views of a static buffer (valid until the next call)."""
        st = self.stack
        if st.buf is None or st._probe() != self._storage_probe:
            raise RuntimeError("the head parameters were moved or replaced after the graph was "
                               "captured; build a new GraphedBinDeltaStep")
        if feat.shape[0] != self.B:
            raise ValueError("GraphedBinDeltaStep was captured for batch %d, got %d" % (self.B, feat.shape[0]))
        self.x.copy_(feat.detach(), non_blocking=True)
        self.label.copy_(label.reshape(-1, 1), non_blocking=True)
        self.bins.copy_(bins.reshape(-1), non_blocking=True)
        self.target.copy_(target.detach().reshape(self.B, -1), non_blocking=True)
        if pose_weight is not None:
            self.weight.fill_(float(pose_weight))
        self.graph.replay()
        if not st.grads_are_mine():          # after zero_grad(set_to_none=True): hand the views out again
            st.publish()
        return self.losses[0], self.losses[1]
