"""Benchmark / profiling legs of the head (BASELINE configs 1, 4, 5): kept out of the product
package; bench.py and profiles/prof_targets.py import them."""
import torch

from bdpose import _lib as L
from bdpose.head import (HeadStack, allreduce_stack_grads, gemm_tf32, onehot, set_precision,
                         sync_head_gradients)


def _pascal_model(C=12, K=200, N0=2048, N1=1000, N2=500, nd=3, seed=0):
    import binDeltaModels as M
    torch.manual_seed(seed)
    m = M.OneBinDeltaModel("none", C, K, N0, N1, N2, nd)
    m.feature_model = torch.nn.Identity()
    return m.cuda()


def _time(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _head_context(dev, m, C, K, hbm, ours):
    """The same Pascal head + loss step (B = 32) through the ORACLE's stock torch modules — on this GPU
    (torch eager: cuBLAS / ATen kernels, the bar SURVEY §2 names) and on the host cores (the
    reference's CPU path) — next to ours.  Benchmark-only use of the oracle."""
    import os
    import time
    import bdpose_oracle as O
    B = 32
    torch.manual_seed(0)
    ref = O.OneBinDeltaHeads(C, K, 2048, 1000, 500, 3)
    x = torch.randn(B, 2048)
    lab = torch.randint(0, C, (B, 1))
    bins = torch.randint(0, K, (B,))
    tgt = torch.randn(B, 3)
    keys = torch.randn(K, 3)

    def step(mod, dev_):
        for p in mod.parameters():
            p.grad = None
        xx = x.to(dev_).requires_grad_(True)
        oh = torch.zeros(B, C, device=dev_).scatter_(1, lab.to(dev_), 1.0)
        y1, y2 = mod(xx, mix=oh)
        lc, lr = O.bin_delta_terms(y1, y2, bins.to(dev_), tgt.to(dev_), keys.to(dev_), "aa")
        (lc + lr).backward()
    # CPU: the reference's path on the host cores (bounded: 1 warm-up + 3 steps)
    torch.set_num_threads(os.cpu_count() or 1)
    ref.train()
    step(ref, "cpu")
    t0 = time.perf_counter()
    for _ in range(3):
        step(ref, "cpu")
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    # stock torch eager on this GPU
    gref = ref.to(dev)
    ms_eager = _time(lambda: step(gref, dev), 20, 10)
    n_params = sum(p.numel() for p in ref.parameters())
    by = 3 * n_params * 4
    return {"torch_eager_gpu": {"ms_fwd_bwd": ms_eager, "samples_per_s_fwd_bwd": B / (ms_eager * 1e-3),
                                "note": "oracle OneBinDeltaHeads (stock nn.Linear / BatchNorm1d) + torch loss "
                                        "ops on the same B200, fp32"},
            "cpu_baseline": {"value": B / (cpu_ms * 1e-3), "unit": "samples/s", "ms_fwd_bwd": cpu_ms,
                             "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": "3 steps of the oracle head + loss (B=32, C=12, K=200) on the host cores"},
            "roofline": {"bound": "hbm", "achieved": by / (ours["ms_fwd_bwd"] * 1e-3) / 1e9, "peak": hbm,
                         "unit": "GB/s", "frac": by / (ours["ms_fwd_bwd"] * 1e-3) / 1e9 / hbm,
                         "algorithmic_bytes_per_step": by,
                         "note": "drop-in eager step through the reference-named modules: weights streamed "
                                 "for fprop, dgrad (reads) and wgrad (write), 3 x %d MB" % (n_params * 4 // 1_000_000)},
            "speedup_vs_torch_eager_gpu": ms_eager / ours["ms_fwd_bwd"]}


def bench(dev, peaks):
    """BASELINE configs 1 and 4: head + fused loss, forward+backward, samples/s."""
    import binDeltaLosses  # noqa: F401  (the fused loss mirrors)
    from bdpose import ops
    out = {}
    hbm = peaks["hbm_gbs"]
    C, K = 12, 200
    m = _pascal_model(C, K).train()
    keys = torch.randn(K, 3, device=dev)
    n_params = sum(p.numel() for p in m.bin_models.parameters()) + sum(p.numel() for p in m.res_models.parameters())
    params = list(m.parameters())
    for B in (32, 96):
        x = torch.randn(B, 2048, device=dev, requires_grad=True)
        lab = torch.randint(0, C, (B, 1), device=dev)
        bins = torch.randint(0, K, (B,), device=dev)
        tgt = torch.randn(B, 3, device=dev)

        def fwd():
            with torch.no_grad():
                m(x, lab)

        def step():
            for p in params:                       # optimizer.zero_grad(set_to_none=True)
                p.grad = None
            y1, y2 = m(x, lab)
            lc, lr, _ = ops.bd_loss(y1, bins, y2, tgt, keys, L.POSE_GEODESIC_AA, True)
            (lc + lr).backward()
        for mode in ("fp32", "tf32"):
            set_precision(mode)
            try:
                ms_f = _time(fwd, 30, 15)
                ms = _time(step, 40, 20)
            finally:
                set_precision("fp32")
            wbytes = n_params * 4
            out["pascal_head_B%d_%s" % (B, mode)] = {
                "samples_per_s_fwd_bwd": B / (ms * 1e-3), "ms_fwd_bwd": ms, "ms_fwd": ms_f,
                "fwd_weight_stream_gbs": wbytes / (ms_f * 1e-3) / 1e9,
                "fwd_hbm_frac": wbytes / (ms_f * 1e-3) / 1e9 / hbm,
                "fwd_bwd_hbm_frac": 3 * wbytes / (ms * 1e-3) / 1e9 / hbm,
                "weight_bytes": wbytes}
    # ---- the bars this leg is measured against ("stock PyTorch eager ops on the same B200"
    #      and the reference's CPU path), config 1: B = 32, fp32 -------------------------------------------
    out["pascal_head_B32_context"] = _head_context(dev, m, C, K, hbm, out["pascal_head_B32_fp32"])
    # opt-in fast path: 10 stacked Parameters instead of 336 per-module ones
    sp = m.stacked_head_parameters()
    for B in (32, 96):
        x = torch.randn(B, 2048, device=dev, requires_grad=True)
        lab = torch.randint(0, C, (B, 1), device=dev)
        bins = torch.randint(0, K, (B,), device=dev)
        tgt = torch.randn(B, 3, device=dev)

        def sstep():
            for p in sp:
                p.grad = None
            y1, y2 = m(x, lab)
            lc, lr, _ = ops.bd_loss(y1, bins, y2, tgt, keys, L.POSE_GEODESIC_AA, True)
            (lc + lr).backward()
        set_precision("tf32")
        try:
            ms = _time(sstep, 40, 20)
        finally:
            set_precision("fp32")
        out["pascal_head_B%d_tf32_stacked_params" % B] = {
            "samples_per_s_fwd_bwd": B / (ms * 1e-3), "ms_fwd_bwd": ms,
            "fwd_bwd_hbm_frac": 3 * n_params * 4 / (ms * 1e-3) / 1e9 / hbm}
    # opt-in CUDA-graph step: heads forward + fused loss + heads backward in one graph launch
    from bdpose.graph_step import GraphedBinDeltaStep
    for B in (32, 96):
        x = torch.randn(B, 2048, device=dev)
        lab = torch.randint(0, C, (B, 1), device=dev)
        bins = torch.randint(0, K, (B,), device=dev)
        tgt = torch.randn(B, 3, device=dev)
        for mode in ("tf32", "fp32"):
            set_precision(mode)
            try:
                gs = GraphedBinDeltaStep(m, B, keys, L.POSE_GEODESIC_AA, True)
                ms = _time(lambda: gs(x, lab, bins, tgt), 50, 20)
            finally:
                set_precision("fp32")
            out["pascal_head_B%d_%s_cuda_graph" % (B, mode)] = {
                "samples_per_s_fwd_bwd": B / (ms * 1e-3), "ms_fwd_bwd": ms,
                "fwd_bwd_hbm_frac": 3 * n_params * 4 / (ms * 1e-3) / 1e9 / hbm}
            del gs
    # config 5: joint category+pose model written the way the reference scripts write it — per-head
    # calls `bin_models[i](x)` mixed with softmax(fc(x)) (learnJointCatPoseModel_weighted.py:107-126,
    # loss 175-180); the head family runs the 2C per-head calls as one fused stack
    mj = _pascal_model(C, K).train()
    fcj = torch.nn.Linear(2048, C).cuda()
    jparams = list(mj.parameters()) + list(fcj.parameters())
    for B in (32, 96):
        x = torch.randn(B, 2048, device=dev, requires_grad=True)
        lab = torch.randint(0, C, (B,), device=dev)
        bins = torch.randint(0, K, (B,), device=dev)
        tgt = torch.randn(B, 3, device=dev)

        def jstep():
            for p in jparams:
                p.grad = None
            y0 = fcj(x)
            mixw = torch.unsqueeze(torch.softmax(y0, dim=1), dim=2)
            y1 = torch.stack([mj.bin_models[i](x) for i in range(C)]).permute(1, 2, 0)
            y2 = torch.stack([mj.res_models[i](x) for i in range(C)]).permute(1, 2, 0)
            y1 = torch.squeeze(torch.bmm(y1, mixw), 2)
            y2 = torch.squeeze(torch.bmm(y2, mixw), 2)
            lc, lr, _ = ops.bd_loss(y1, bins, y2, tgt, keys, L.POSE_GEODESIC_AA, True)
            (0.1 * torch.nn.functional.cross_entropy(y0, lab) + lc + lr).backward()
        set_precision("tf32")
        try:
            ms = _time(jstep, 30, 15)
        finally:
            set_precision("fp32")
        out["joint_weighted_script_style_B%d_tf32" % B] = {
            "samples_per_s_fwd_bwd": B / (ms * 1e-3), "ms_fwd_bwd": ms}
    del mj
    # raw fc1 GEMM: the dominant kernel of the head (197 MB of weights streamed once)
    H, N1, N0, B = 24, 1000, 2048, 32
    w1 = m._heads().ensure()["w1"]
    xb = torch.randn(B, N0, device=dev)
    h1 = torch.empty(B, H * N1, device=dev)
    for precise in (True, False):
        ms = _time(lambda: gemm_tf32(xb, 0, N0, 0, w1, 0, N0, 0, h1, 0, H * N1, 0, B, H * N1, N0, precise=precise), 30, 5)
        by = w1.numel() * 4 + xb.numel() * 4 + h1.numel() * 4
        out["fc1_gemm_B32_%s" % ("3xtf32" if precise else "tf32")] = {
            "ms": ms, "achieved_gbs": by / (ms * 1e-3) / 1e9, "hbm_frac": by / (ms * 1e-3) / 1e9 / hbm,
            "tflops": 2.0 * H * N1 * N0 * B / (ms * 1e-3) / 1e12}
    del m
    # config 4: ObjectNet one-hot-concat heads, C=100, B=256
    import objectnetHelperFunctions as OH
    torch.manual_seed(0)
    om = OH.OneBinDeltaModel.__new__(OH.OneBinDeltaModel)
    torch.nn.Module.__init__(om)
    om.num_classes, om.num_clusters = 100, 200
    om.feature_model = torch.nn.Identity()
    om.bin_model = OH.bin_3layer(2148, 1000, 500, 200).cuda()
    om.res_model = OH.res_3layer(2148, 1000, 500, 3).cuda()
    object.__setattr__(om, "_stack", None)
    om.train()
    B = 256
    x = torch.randn(B, 2048, device=dev, requires_grad=True)
    lab = torch.randint(0, 100, (B, 1), device=dev)
    bins = torch.randint(0, 200, (B,), device=dev)
    tgt = torch.randn(B, 3, device=dev)
    keys = torch.randn(200, 3, device=dev)

    oparams = list(om.parameters())

    def ostep():
        for p in oparams:
            p.grad = None
        y1, y2 = om.forward_features(x, lab)
        lc, lr, _ = ops.bd_loss(y1, bins, y2, tgt, keys, L.POSE_GEODESIC_AA, True)
        (lc + 10 * lr).backward()
    for mode in ("fp32", "tf32"):
        set_precision(mode)
        try:
            ms = _time(ostep, 40, 20)
        finally:
            set_precision("fp32")
        out["objectnet_head_B256_%s" % mode] = {"samples_per_s_fwd_bwd": B / (ms * 1e-3), "ms_fwd_bwd": ms}
    return out


def bench_dp(dev, world):
    """BASELINE config 4 / north star (e): batch data-parallel head step on every rank.  Each rank
    runs the CUDA-graph step (heads forward, fused loss, heads backward: one graph launch) on its own
    batch (per-GPU BatchNorm statistics), then ONE NCCL all-reduce(AVG) over the flat gradient
    buffer.  The eager autograd step is measured next to it: with 8 processes on one host it is
    host-bound (the Python issue path of every rank competes for the same cores).
    Returns aggregate samples/s (time = max over ranks)."""
    import torch.distributed as dist
    import objectnetHelperFunctions as OH
    from bdpose import ops
    from bdpose.graph_step import GraphedBinDeltaStep
    out = {}

    def timed(step):
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 30], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    set_precision("tf32")
    try:
        for name in ("objectnet", "pascal"):
            torch.manual_seed(0)
            keys = torch.randn(200, 3, device=dev)
            if name == "objectnet":
                m = OH.OneBinDeltaModel.__new__(OH.OneBinDeltaModel)
                torch.nn.Module.__init__(m)
                m.num_classes, m.num_clusters = 100, 200
                m.feature_model = torch.nn.Identity()
                m.bin_model = OH.bin_3layer(2148, 1000, 500, 200).cuda()
                m.res_model = OH.res_3layer(2148, 1000, 500, 3).cuda()
                object.__setattr__(m, "_stack", None)
                B, C = 256, 100
                stack = HeadStack([[m.bin_model], [m.res_model]])
                object.__setattr__(m, "_stack", stack)
            else:
                m = _pascal_model()
                B, C = 96, 12
                stack = m._heads()
            m.train()
            params = list(m.parameters())
            x = torch.randn(B, 2048, device=dev, requires_grad=True)
            lab = torch.randint(0, C, (B, 1), device=dev)
            bins = torch.randint(0, 200, (B,), device=dev)
            tgt = torch.randn(B, 3, device=dev)

            def eager():
                for p in params:
                    p.grad = None
                y1, y2 = m.forward_features(x, lab)
                lc, lr, _ = ops.bd_loss(y1, bins, y2, tgt, keys, L.POSE_GEODESIC_AA, True)
                (lc + lr).backward()
                sync_head_gradients(m)
            ms_eager = timed(eager)
            gs = GraphedBinDeltaStep(stack, B, keys, L.POSE_GEODESIC_AA, True)
            zero_lab = torch.zeros_like(lab)

            def graphed():
                if name == "objectnet":       # one MLP pair for all classes on cat(features, one-hot)
                    gs(torch.cat((x.detach(), onehot(lab, C)), dim=1), zero_lab, bins, tgt)
                else:
                    gs(x, lab, bins, tgt)
                allreduce_stack_grads(stack)
            ms = timed(graphed)
            n_par = sum(p.numel() for p in params)
            out["%s_head_dp%d_B%d_tf32" % (name, world, B)] = {
                "samples_per_s_fwd_bwd": B * world / (ms * 1e-3), "ms_fwd_bwd_allreduce": ms,
                "ms_eager_autograd_step": ms_eager, "step": "cuda graph + one NCCL all-reduce(AVG)",
                "allreduce_bytes": n_par * 4, "per_gpu_batch": B}
            del m, gs
    finally:
        set_precision("fp32")
    return out


def profile(dev):
    """One training step of the Pascal head per precision mode (tf32, then 3xTF32) for ncu
    (profiles/prof_targets.py head)."""
    from bdpose import ops
    m = _pascal_model().train()
    B, K = 32, 200
    x = torch.randn(B, 2048, device=dev, requires_grad=True)
    lab = torch.randint(0, 12, (B, 1), device=dev)
    bins = torch.randint(0, K, (B,), device=dev)
    tgt = torch.randn(B, 3, device=dev)
    keys = torch.randn(K, 3, device=dev)
    try:
        for mode in ("tf32", "fp32"):
            set_precision(mode)
            y1, y2 = m(x, lab)
            lc, lr, _ = ops.bd_loss(y1, bins, y2, tgt, keys, L.POSE_GEODESIC_AA, True)
            (lc + lr).backward()
    finally:
        set_precision("fp32")
