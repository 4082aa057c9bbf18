import os, sys, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200"))
import torch
import bench_head
from bdpose import head, ops, _lib as L
dev = torch.device("cuda", 0)
C, K, B = 12, 200, 32
mj = bench_head._pascal_model(C, K).train()
fcj = torch.nn.Linear(2048, C).cuda()
jparams = list(mj.parameters()) + list(fcj.parameters())
keys = torch.randn(K, 3, device=dev)
x = torch.randn(B, 2048, device=dev, requires_grad=True)
lab = torch.randint(0, C, (B,), device=dev); bins = torch.randint(0, K, (B,), device=dev); tgt = torch.randn(B, 3, device=dev)
head.set_precision("tf32")
def jstep():
    for p in jparams: p.grad = None
    y0 = fcj(x)
    mixw = torch.unsqueeze(torch.softmax(y0, dim=1), dim=2)
    y1 = torch.stack([mj.bin_models[i](x) for i in range(C)]).permute(1, 2, 0)
    y2 = torch.stack([mj.res_models[i](x) for i in range(C)]).permute(1, 2, 0)
    y1 = torch.squeeze(torch.bmm(y1, mixw), 2); y2 = torch.squeeze(torch.bmm(y2, mixw), 2)
    lc, lr, _ = ops.bd_loss(y1, bins, y2, tgt, keys, L.POSE_GEODESIC_AA, True)
    (0.1 * torch.nn.functional.cross_entropy(y0, lab) + lc + lr).backward()
for _ in range(10): jstep()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(30): jstep()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
