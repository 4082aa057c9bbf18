"""Round-2 head parity: the ObjectNet one-hot-concat head and the reference's
JointCatPoseModel re-parenting pattern (§8b "script-defined forwards"), against

  * golden vectors of the REFERENCE's classes (tests/golden/make_golden.py objectnet / joint:
    objectnetHelperFunctions.OneBinDeltaModel; learnJointCatPoseModel_weighted.JointCatPoseModel
    compiled from the script's own class statement, both `multires` branches), and
  * the oracle at full size (C=100, n0=2048+100, K=200, B=256 and 96).

Tolerances as tests/test_gpu_head.py: forward 1e-5 of the tensor scale in "fp32" head precision,
2e-3 in "tf32"; golden gradients 1e-4; full-size gradients robustly (flip_close)."""
import copy

import numpy as np
import pytest
import torch

import bdpose_oracle as O
from test_gpu_head import scale_close, flip_close, FP32_TOL, TF32_TOL, GRAD_TOL

pytestmark = pytest.mark.gpu


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith(prefix)}


def _objectnet_model(C, K, n0, n1, n2, nd):
    import objectnetHelperFunctions as OH
    m = OH.OneBinDeltaModel.__new__(OH.OneBinDeltaModel)
    torch.nn.Module.__init__(m)
    m.num_classes, m.num_clusters = C, K
    m.feature_model = torch.nn.Identity()
    m.bin_model = OH.bin_3layer(n0 + C, n1, n2, K).cuda()
    m.res_model = OH.res_3layer(n0 + C, n1, n2, nd).cuda()
    object.__setattr__(m, "_stack", None)
    return m


def test_objectnet_head_golden(cuda, golden):
    """objectnetHelperFunctions.OneBinDeltaModel (155-172) of the reference: train forward, every
    gradient, BatchNorm running statistics, eval forward."""
    g = golden("objectnet_head")
    C, K, n0, n1, n2, nd, B = [int(v) for v in g["dims"]]
    m = _objectnet_model(C, K, n0, n1, n2, nd)
    m.load_state_dict(_sd(g, "sd0/"))
    m.cuda().train()
    x = torch.from_numpy(g["x"]).to(cuda).requires_grad_(True)
    lab = torch.from_numpy(g["label"]).to(cuda)
    y1, y2 = m(x, lab)
    ((y1 * torch.from_numpy(g["w1"]).to(cuda)).sum() + (y2 * torch.from_numpy(g["w2"]).to(cuda)).sum()).backward()
    scale_close(y1, torch.from_numpy(g["train_y1"]), FP32_TOL, "objectnet golden y1")
    scale_close(y2, torch.from_numpy(g["train_y2"]), FP32_TOL, "objectnet golden y2")
    scale_close(x.grad, torch.from_numpy(g["train_gx"]), GRAD_TOL, "objectnet golden dx")
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        scale_close(p.grad, torch.from_numpy(g["train_grad/" + k]), GRAD_TOL, "objectnet golden grad " + k)
    sd = m.state_dict()
    for k in g.files:
        if k.startswith("train_sd/"):
            name = k[len("train_sd/"):]
            if "num_batches" in name:
                assert int(sd[name]) == int(g[k]), name
            else:
                scale_close(sd[name], torch.from_numpy(g[k]), FP32_TOL, "objectnet golden " + name)
    m.eval()
    with torch.no_grad():
        e1, e2 = m(x.detach(), lab)
    scale_close(e1, torch.from_numpy(g["eval_y1"]), FP32_TOL, "objectnet golden eval y1")
    scale_close(e2, torch.from_numpy(g["eval_y2"]), FP32_TOL, "objectnet golden eval y2")


@pytest.mark.parametrize("B", [256, 96])
def test_objectnet_head_vs_oracle_full_size(cuda, B):
    """BASELINE config 4 shapes: C=100 one-hot concat (2148 input columns: the TMA row pitch is not
    a multiple of 128 bytes), K=200, 2148-1000-500, B=256 / 96 (the script's batch,
    learnObjectnetBDModel.py:75).  fp32 mode against the oracle in float64 and float32; tf32 mode
    forward at 2e-3; train + eval; all gradients; BatchNorm statistics."""
    from bdpose import head
    torch.manual_seed(2)
    C, K, n0, n1, n2, nd = 100, 200, 2048, 1000, 500, 3
    ref = O.ObjectnetHeads(C, K, n0, n1, n2, nd)
    ref64 = copy.deepcopy(ref).double()
    m = _objectnet_model(C, K, n0, n1, n2, nd)
    m.load_state_dict(ref.state_dict())
    m.cuda().train(); ref.train(); ref64.train()
    x = torch.randn(B, n0)
    lab = torch.randint(0, C, (B, 1))
    w1, w2 = torch.randn(B, K), torch.randn(B, nd)

    def step(mod, dev, dt):
        xs = x.detach().clone().to(dev, dt).requires_grad_(True)
        y1, y2 = mod(xs, lab.to(dev))
        ((y1 * w1.to(dev, dt)).sum() + (y2 * w2.to(dev, dt)).sum()).backward()
        return y1, y2, xs
    r1, r2, rx = step(ref, "cpu", torch.float32)
    d1, d2, dx = step(ref64, "cpu", torch.float64)
    y1, y2, gx = step(m, cuda, torch.float32)
    scale_close(y1, d1, FP32_TOL, "objectnet y1 vs f64"); scale_close(y2, d2, FP32_TOL, "objectnet y2 vs f64")
    scale_close(y1, r1, FP32_TOL, "objectnet y1 vs f32"); scale_close(y2, r2, FP32_TOL, "objectnet y2 vs f32")
    flip_close(gx.grad, dx.grad, "objectnet dx vs f64")
    flip_close(gx.grad, rx.grad, "objectnet dx vs f32")
    p32, p64 = dict(ref.named_parameters()), dict(ref64.named_parameters())
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        flip_close(p.grad, p64[k].grad, "objectnet %s vs f64" % k)
        flip_close(p.grad, p32[k].grad, "objectnet %s vs f32" % k)
    rsd = ref.state_dict()
    for k, v in m.state_dict().items():
        if "running" in k:
            scale_close(v, rsd[k], FP32_TOL, "objectnet " + k)
        elif "num_batches" in k:
            assert int(v) == int(rsd[k]) == 1
    # eval mode (running statistics), fp32 and tf32 head precision
    m.eval(); ref64.eval()
    with torch.no_grad():
        q1, q2 = ref64(x.double(), lab)
        e1, e2 = m(x.to(cuda), lab.to(cuda))
        scale_close(e1, q1, FP32_TOL, "objectnet eval y1"); scale_close(e2, q2, FP32_TOL, "objectnet eval y2")
        head.set_precision("tf32")
        try:
            t1, t2 = m(x.to(cuda), lab.to(cuda))
        finally:
            head.set_precision("fp32")
        scale_close(t1, q1, TF32_TOL, "objectnet tf32 y1"); scale_close(t2, q2, TF32_TOL, "objectnet tf32 y2")
    # tf32 train-mode forward against the float64 oracle (same BatchNorm batch statistics)
    m.train(); ref64.train()
    head.set_precision("tf32")
    try:
        with torch.no_grad():
            t1, t2 = m(x.to(cuda), lab.to(cuda))
            q1, q2 = ref64(x.double(), lab)
    finally:
        head.set_precision("fp32")
    scale_close(t1, q1, TF32_TOL, "objectnet tf32 train y1"); scale_close(t2, q2, TF32_TOL, "objectnet tf32 train y2")


class _ScriptJoint(torch.nn.Module):
    """The call pattern of learnJointCatPoseModel_weighted.JointCatPoseModel (94-126), written against
    the public attributes it uses: lifts feature_model / bin_models / res_models out of a bin-delta
    model into a new parent, adds `fc`, and calls the heads ONE BY ONE inside forward.  The golden
    outputs it is compared with come from the reference's own class statement."""

    def __init__(self, oracle_model, N0, multires):
        super().__init__()
        self.num_classes = oracle_model.num_classes
        self.num_clusters = oracle_model.num_clusters
        self.ndim = oracle_model.ndim
        self.feature_model = oracle_model.feature_model
        self.bin_models = oracle_model.bin_models
        self.res_models = oracle_model.res_models
        self.fc = torch.nn.Linear(N0, self.num_classes).cuda()
        self.multires = multires

    def forward(self, x):
        x = self.feature_model(x)
        y0 = self.fc(x)
        label = torch.unsqueeze(torch.softmax(y0, dim=1), dim=2)
        y1 = torch.stack([self.bin_models[i](x) for i in range(self.num_classes)]).permute(1, 2, 0)
        y1 = torch.squeeze(torch.bmm(y1, label), 2)
        if not self.multires:
            y2 = torch.stack([self.res_models[i](x) for i in range(self.num_classes)]).permute(1, 2, 0)
            y2 = torch.squeeze(torch.bmm(y2, label), 2)
        else:
            y2 = torch.stack([self.res_models[i](x) for i in range(self.num_classes * self.num_clusters)])
            y2 = y2.view(self.num_classes, self.num_clusters, -1, self.ndim).permute(1, 2, 3, 0)
            y2 = torch.squeeze(torch.matmul(y2, label), 3)
            pose_label = torch.argmax(y1, dim=1, keepdim=True)
            pose_label = torch.zeros(pose_label.size(0), self.num_clusters, device=x.device).scatter_(1, pose_label, 1.0)
            y2 = torch.squeeze(torch.bmm(y2.permute(1, 2, 0), pose_label.unsqueeze(2)), 2)
        return [y0, y1, y2]


@pytest.mark.parametrize("multires", [False, True])
def test_joint_cat_pose_model_golden(cuda, golden, multires):
    """Re-parenting + per-head calls: outputs, every gradient and the BatchNorm
    statistics equal those of the reference's JointCatPoseModel over the reference's models, and the
    per-head calls of one forward share fused launches (one run per sibling list)."""
    import binDeltaModels as M
    from bdpose import head
    g = golden("joint")
    C, K, N0, N1, N2, N3, nd, B = [int(v) for v in g["dims"]]
    tag = "mr/" if multires else "sr/"
    if multires:
        base = M.OneDeltaPerBinModel("none", C, K, N0, N1, N2, N3, nd)
    else:
        base = M.OneBinDeltaModel("none", C, K, N0, N1, N2, nd)
    base.feature_model = torch.nn.Identity()
    model = _ScriptJoint(base, N0, multires)
    missing = model.load_state_dict(_sd(g, tag + "sd0/"), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.cuda().train()
    calls = {"all": 0, "mlp2": 0}
    orig_all, orig_mlp2 = head.run_heads_all, head.run_mlp2_all

    def count_all(*a, **k):
        calls["all"] += 1
        return orig_all(*a, **k)

    def count_mlp2(*a, **k):
        calls["mlp2"] += 1
        return orig_mlp2(*a, **k)
    head.run_heads_all, head.run_mlp2_all = count_all, count_mlp2
    try:
        x = torch.from_numpy(g[tag + "x"]).to(cuda).requires_grad_(True)
        ys = model(x)
        assert calls["all"] == 1 and calls["mlp2"] == (1 if multires else 0), calls
        sum((y * torch.from_numpy(g[tag + "w%d" % i]).to(cuda)).sum() for i, y in enumerate(ys)).backward()
    finally:
        head.run_heads_all, head.run_mlp2_all = orig_all, orig_mlp2
    for i, y in enumerate(ys):
        scale_close(y, torch.from_numpy(g[tag + "train_y%d" % i]), FP32_TOL, "joint %s y%d" % (tag, i))
    scale_close(x.grad, torch.from_numpy(g[tag + "train_gx"]), GRAD_TOL, "joint %s dx" % tag)
    for k, p in model.named_parameters():
        ref = torch.from_numpy(g[tag + "train_grad/" + k])
        assert p.grad is not None, k
        scale_close(p.grad, ref, GRAD_TOL, "joint %s grad %s" % (tag, k))
    sd = model.state_dict()
    for k in g.files:
        if k.startswith(tag + "train_sd/"):
            name = k[len(tag + "train_sd/"):]
            if "num_batches" in name:
                assert int(sd[name]) == int(g[k]), name
            else:
                scale_close(sd[name], torch.from_numpy(g[k]), FP32_TOL, "joint %s %s" % (tag, name))
    model.eval()
    with torch.no_grad():
        es = model(x.detach())
    for i, y in enumerate(es):
        scale_close(y, torch.from_numpy(g[tag + "eval_y%d" % i]), FP32_TOL, "joint %s eval y%d" % (tag, i))


def test_weights_changed_between_forward_and_backward(cuda):
    """Stock torch refuses a backward whose saved weights were modified in place; the fused heads read
    the current stacked weights in backward, so an optimizer step in between must be refused too
    (sentinel version counters) — and the normal two-forwards-one-backward pattern of
    the scripts (learnGeodesicBDModel.py:116-120, 183) must keep working."""
    import binDeltaModels as M
    torch.manual_seed(3)
    m = M.OneBinDeltaModel("none", 3, 16, 64, 40, 24, 3)
    m.feature_model = torch.nn.Identity()
    m.cuda().train()
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    x = torch.randn(8, 64, device=cuda)
    lab = torch.randint(0, 3, (8, 1), device=cuda)
    y1, y2 = m(x, lab)
    z1, z2 = m(x * 0.5, lab)
    (y1.sum() + y2.sum() + z1.sum() + z2.sum()).backward()      # two forwards, one backward: fine
    opt.step()
    opt.zero_grad()
    y1, y2 = m(x, lab)
    opt.zero_grad()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(1e-3)                                        # what an optimizer step does
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        (y1.sum() + y2.sum()).backward()
