import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-regression_b200")); sys.path.insert(0, ROOT)
import torch
import bench_head, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
def t(fn, n=20, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / n], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)
for n in (1_000_000, 5_403_703, 20_000_000, 62_444_436):
    x = torch.randn(n, device=dev)
    a = t(lambda: dist.all_reduce(x, op=dist.ReduceOp.AVG))
    s = t(lambda: dist.all_reduce(x, op=dist.ReduceOp.SUM))
    if rank == 0: print("allreduce %9d floats: AVG %.3f ms  SUM %.3f ms" % (n, a, s), flush=True)
from bdpose import head
r = bench_head.bench_dp(dev, world)
if rank == 0:
    for k, v in r.items(): print(k, v, flush=True)
# the same step without the collective
import objectnetHelperFunctions as OH
from bdpose import ops, _lib as L
head.set_precision("tf32")
m = OH.OneBinDeltaModel.__new__(OH.OneBinDeltaModel); torch.nn.Module.__init__(m)
m.num_classes, m.num_clusters = 100, 200; m.feature_model = torch.nn.Identity()
m.bin_model = OH.bin_3layer(2148, 1000, 500, 200).cuda(); m.res_model = OH.res_3layer(2148, 1000, 500, 3).cuda()
object.__setattr__(m, "_stack", None); m.train()
params = list(m.parameters()); B = 256
x = torch.randn(B, 2048, device=dev, requires_grad=True); lab = torch.randint(0, 100, (B, 1), device=dev)
bins = torch.randint(0, 200, (B,), device=dev); tgt = torch.randn(B, 3, device=dev); keys = torch.randn(200, 3, device=dev)
def step(sync):
    for p in params: p.grad = None
    y1, y2 = m.forward_features(x, lab)
    lc, lr, _ = ops.bd_loss(y1, bins, y2, tgt, keys, L.POSE_GEODESIC_AA, True)
    (lc + lr).backward()
    if sync: head.sync_head_gradients(m)
a = t(lambda: step(False)); b = t(lambda: step(True))
if rank == 0: print("objectnet step: no sync %.3f ms, with sync %.3f ms" % (a, b), flush=True)
dist.barrier(); dist.destroy_process_group()
