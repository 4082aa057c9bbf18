import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "multi-modal-regression_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import bdpose_oracle as O
import binDeltaModels as M
torch.manual_seed(0)
C, K, N0, N1, N2, nd, B = 12, 200, 2048, 1000, 500, 3, 32
ref = O.OneBinDeltaHeads(C, K, N0, N1, N2, nd)
model = M.OneBinDeltaModel("none", C, K, N0, N1, N2, nd)
model.feature_model = torch.nn.Identity()
model.load_state_dict(ref.state_dict())
model.cuda().train(); ref.train()
ref64 = copy.deepcopy(ref).double()
for trial in range(3):
    x = torch.randn(B, N0); lab = torch.randint(0, C, (B, 1))
    w1, w2 = torch.randn(B, K), torch.randn(B, nd)
    res = {}
    for name, m, dev, dt in (("cpu32", ref, "cpu", torch.float32), ("cpu64", ref64, "cpu", torch.float64), ("gpu", model, "cuda", torch.float32)):
        xx = x.detach().clone().to(dev, dt).requires_grad_(True)
        for p in m.parameters(): p.grad = None
        y1, y2 = m(xx, label=lab.to(dev)) if name != "gpu" else m(xx, lab.to(dev))
        ((y1 * w1.to(dev, dt)).sum() + (y2 * w2.to(dev, dt)).sum()).backward()
        res[name] = (y1.detach().double().cpu(), xx.grad.double().cpu(), dict((k, p.grad.double().cpu()) for k, p in m.named_parameters()))
    def rel(a, b):
        e = (a - b).abs(); s = b.abs().max()
        return "med %.1e max %.1e" % (float(e.median() / s), float(e.max() / s))
    print("trial", trial)
    print("  y1   cpu32 vs 64:", rel(res["cpu32"][0], res["cpu64"][0]), "| gpu vs 64:", rel(res["gpu"][0], res["cpu64"][0]))
    print("  dx   cpu32 vs 64:", rel(res["cpu32"][1], res["cpu64"][1]), "| gpu vs 64:", rel(res["gpu"][1], res["cpu64"][1]))
    for k in ("bin_models.0.fc1.weight", "bin_models.5.fc2.weight", "res_models.3.fc1.weight", "bin_models.2.bn1.weight"):
        print("  %-26s cpu32 vs 64:" % k, rel(res["cpu32"][2][k], res["cpu64"][2][k]), "| gpu vs 64:", rel(res["gpu"][2][k], res["cpu64"][2][k]))
