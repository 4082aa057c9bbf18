import torch, time
dev = torch.device("cuda", 0)
n = 120_000_000 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(200_000_000 // 4, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device=dev)
d_out = torch.empty(200_000_000 // 4, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, k=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print("H2D 120MB %.2f ms (%.1f GB/s)  D2H 200MB %.2f ms (%.1f GB/s)  concurrent %.2f ms" % (a, 0.12 / a * 1e3, b, 0.2 / b * 1e3, c))
