#!/usr/bin/env python
"""bench.py — headline benchmark of the bin-and-delta pose hot path on B200.

Workload (BASELINE.json configs[1]): binDeltaGenerators label generation — 10 M synthetic rotations
(axis-angle, fp32) against a K=1000 pose dictionary: nearest key (fp64-faithful argmin, int64 bin) +
residual delta (fp32).  One step = one pass over the 10 M rotations of a rank.  With --gpus N each
rank labels its own 10 M rotations (weak scaling, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extras]

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` goes through the public
API (binDeltaGenerators.assign_labels) from pinned host buffers and back, `roofline` is the assign
kernel against the measured HBM peak, `cpu_baseline` is the reference's CPU path (scikit-learn
KMeans.predict + numpy residual, the call binDeltaGenerators.py:27-30 makes) on a bounded sample.
`extras` carries the other configs of BASELINE.json (k-means Lloyd, fused loss, evaluation, head).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "multi-modal-regression_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_ROT = 10_000_000
K_DICT = 1000
BYTES_PER_ROT = 12 + 8 + 12          # fp32 y in, int64 bin out, fp32 residual out
METRIC = "label-generation rotations/s (nearest key + residual delta, 10M rotations, K=1000)"


def synth_rotations(n, seed, device):
    """Uniform rotations on SO(3) as axis-angle fp32 [n,3]."""
    g = torch.Generator(device=device).manual_seed(seed)
    q = torch.randn(n, 4, device=device, generator=g, dtype=torch.float32)
    q = q / q.norm(dim=1, keepdim=True)
    w = q[:, :1].abs().clamp(max=1)
    v = q[:, 1:] * torch.sign(q[:, :1])
    return (v / v.norm(dim=1, keepdim=True).clamp_min(1e-30) * (2 * torch.acos(w))).contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ---------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own call (sklearn predict + numpy residual), bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_label_generation(y, centers, repeats=1):
    """Returns (rotations/s, kind, cores).  y [n,3] float32 numpy, centers [K,3] float64."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    try:
        from sklearn.cluster import KMeans
        km = KMeans(n_clusters=centers.shape[0], init=centers, n_init=1, max_iter=1)
        km.cluster_centers_ = np.ascontiguousarray(centers)
        km._n_threads = cores
        km.n_features_in_ = centers.shape[1]
        km._n_init = 1

        def run():
            b = km.predict(y.astype(np.float64))                  # binDeltaGenerators.py:27
            return b, (y - centers[b, :]).astype(np.float32)      # :30-31
        kind = "reference"
    except Exception:
        import bdpose_oracle as O

        def run():
            return O.predict_residual(y, centers)
        kind = "port"
    run()
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        run()
        best = min(best, time.perf_counter() - t0)
    return y.shape[0] / best, kind, cores


def run_reference(args, rank, world):
    """--impl reference: the CPU path alone, on rank 0 only."""
    if rank != 0:
        return
    n = 2_000_000
    y = synth_rotations(n + K_DICT, 0, "cpu").numpy()
    centers = y[:K_DICT].astype(np.float64)
    y = y[K_DICT:]
    for _ in range(max(args.warmup, 1)):
        cpu_label_generation(y[:200_000], centers)
    t0 = time.perf_counter()
    kind = "port"
    cores = os.cpu_count() or 1
    for _ in range(args.steps):
        _, kind, cores = cpu_label_generation(y, centers, repeats=1)
    # cpu_label_generation runs one untimed + one timed pass per call; time the whole loop honestly
    dt = (time.perf_counter() - t0) / (2 * args.steps)
    val = n / dt
    sample = "%d rotations per step against K=%d (of the 10M workload), host cores" % (n, K_DICT)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "rotations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "configs[1]: label generation, 10M rotations, K=1000 (bounded sample)",
                       "n_rotations": n, "K": K_DICT},
            "cpu_baseline": {"value": val, "unit": "rotations/s", "cores": cores, "kind": kind,
                             "sample": sample},
            "e2e": {"value": val, "unit": "rotations/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# extras: the other BASELINE.json configs, short runs
# ---------------------------------------------------------------------------------------------------
def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps       # ms


def extras(dev, x, centers, peaks, rank, world):
    from bdpose import ops, kmeans, _lib as L
    import torch.distributed as dist
    out = {}
    hbm = peaks["hbm_gbs"]
    # ---- config 3: k-means Lloyd iterations, 10 M rotations TOTAL sharded over the ranks --------
    n_local = N_ROT // world
    xs = x[:n_local].double().contiguous()
    for K in (200, 1000):
        init = centers[:K].clone()

        def fit_ms(iters):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            kmeans.kmeans_lloyd(xs, init, fixed_iters=iters, group=None)
            t1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms)
        fit_ms(2)                                                       # warm-up
        short, long_ = fit_ms(2), fit_ms(22)
        # marginal cost of one Lloyd iteration (E+M step, grid rebuild, all-reduce, finalisation):
        # the fit's one-off work (centring, variance, final E-step) cancels in the difference
        per_iter = (long_ - short) / 20.0
        out["kmeans_K%d" % K] = {"rotation_iterations_per_s": N_ROT / (per_iter * 1e-3),
                                 "ms_per_iteration": per_iter, "fit_ms_22_iterations": long_,
                                 "n_rotations_total": N_ROT, "scaling": "strong",
                                 "hbm_frac": (n_local * 28.0 / (per_iter * 1e-3)) / 1e9 / hbm}
    # ---- config 4: data-parallel head step (all ranks), gradients all-reduced over NCCL ----------
    if world > 1:
        from bdpose import head as _head
        out.update(_head.bench_dp(dev, world))
    if rank != 0:
        return out
    # ---- config 1b: fused loss fwd+bwd, 1 M rows, K=200 ------------------------------------------
    B, K = 1_000_000, 200
    score = torch.randn(B, K, device=dev)
    bins = torch.randint(0, K, (B,), device=dev)
    delta = torch.randn(B, 3, device=dev) * 0.2
    target = x[:B].contiguous()
    keys = centers[:K].float().contiguous()
    ms = timed(lambda: ops.bd_loss_raw(score, bins, delta, target, keys, L.POSE_GEODESIC_AA, True), 10, 3)
    by = B * (2 * K * 4 + 8 + 12 + 12 + 12 + 8)
    out["fused_loss_1M_K200"] = {"samples_per_s": B / (ms * 1e-3), "ms": ms, "bytes": by,
                                 "achieved_gbs": by / (ms * 1e-3) / 1e9, "hbm_frac": by / (ms * 1e-3) / 1e9 / hbm}
    del score
    # small-batch training shape (launch-latency bound)
    B2 = 96
    s2 = torch.randn(B2, K, device=dev); b2 = bins[:B2]; d2 = delta[:B2].contiguous(); t2 = target[:B2].contiguous()
    ms = timed(lambda: ops.bd_loss_raw(s2, b2, d2, t2, keys, L.POSE_GEODESIC_AA, True), 50, 5)
    out["fused_loss_B96_K200"] = {"samples_per_s": B2 / (ms * 1e-3), "us_per_step": ms * 1e3}
    # ---- config 5: evaluation over 1 M predictions -------------------------------------------------
    a = x[:1_000_000].contiguous(); b = x[1_000_000:2_000_000].contiguous()
    labels = torch.randint(0, 12, (1_000_000,), device=dev)

    def ev():
        e = ops.geodesic_error_deg(a, b)
        ops.error_stats(e, labels, 12)
    ms = timed(ev, 10, 3)
    out["eval_1M"] = {"pairs_per_s": 1e6 / (ms * 1e-3), "ms": ms}
    ms = timed(lambda: ops.geodesic_error_deg(a, b), 10, 3)
    by = 1_000_000 * (12 + 12 + 8)
    out["eval_1M"]["error_kernel_ms"] = ms
    out["eval_1M"]["error_kernel_hbm_frac"] = by / (ms * 1e-3) / 1e9 / hbm
    # ---- config 1 / 4: head fwd+bwd, when the head kernels are built -------------------------------
    try:
        from bdpose import head
        if hasattr(head, "bench"):
            out.update(head.bench(dev, peaks))
    except ImportError:
        pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--brute", action="store_true", help="also time the brute-force N*K scan")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 prints exactly one line on stdout: NCCL's version banner (written by the library
        # itself at communicator creation) goes to stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    from bdpose import ops, _lib as L
    import binDeltaGenerators as G
    L.lib()
    peaks, peak_src = measured_peaks()

    # synthetic workload of this rank: 10 M rotations + a K=1000 dictionary (seeded)
    x = synth_rotations(N_ROT, 1000 + rank, dev)
    centers = synth_rotations(K_DICT, 7, dev).double().contiguous()

    def step():
        # public op: builds the key grid of the dictionary (3 small launches) + one pruned query launch
        return ops.assign_nearest(x, centers, want_residual=True)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    clocks = clk.summary()
    if clocks["samples"] < 3:
        # the timed region is only a few ms: keep the same step running for ~0.7 s so that
        # nvidia-smi sees the clocks and throttle reasons under this load
        with ClockSampler(local) as clk2:
            t_end = time.perf_counter() + 0.7
            while time.perf_counter() < t_end:
                for _ in range(20):
                    step()
                torch.cuda.synchronize()
        rows = clk.rows + clk2.rows
        clk2.rows = rows
        clocks = clk2.summary()
        clocks["note"] = "timed region is %.1f ms; sampled over it plus 0.7 s more of the same steps" % float(ms_total)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total) / args.steps
    value = N_ROT * world / (ms_step * 1e-3)
    # dominant kernel alone (the pruned query against a prebuilt grid), CUDA events on its stream
    grid = ops.KeyGrid(centers)
    for _ in range(3):
        ops.assign_nearest(x, centers, grid=grid)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        ops.assign_nearest(x, centers, grid=grid)
    e1.record()
    torch.cuda.synchronize()
    kernel_s = e0.elapsed_time(e1) / args.steps * 1e-3
    achieved = N_ROT * BYTES_PER_ROT / kernel_s / 1e9
    brute_ms = None
    if args.brute:
        # the brute-force N*K scan (same labels; the reference's algorithm) for comparison
        ops.assign_nearest(x, centers, grid=None)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            ops.assign_nearest(x, centers, grid=None)
        e1.record()
        torch.cuda.synchronize()
        brute_ms = e0.elapsed_time(e1) / 3

    # ---- e2e: public API from pinned host buffers, results read back to pinned host buffers -----
    y_host = x.cpu().pin_memory()
    bin_host = torch.empty(N_ROT, dtype=torch.int64).pin_memory()
    res_host = torch.empty(N_ROT, 3, dtype=torch.float32).pin_memory()

    def e2e_step():
        # public host-to-host API: chunked H2D -> pruned query -> D2H over three streams
        G.assign_labels_host(y_host, centers, out_bin=bin_host, out_res=res_host)
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    k_e2e = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(k_e2e):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = torch.tensor([e0.elapsed_time(e1) / k_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_val = N_ROT * world / (float(ms_e2e) * 1e-3)
    del y_host, bin_host, res_host

    ex = {}
    if not args.no_extras:
        ex = extras(dev, x, centers, peaks, rank, world)

    if rank == 0:
        # CPU baseline on a bounded sample of the same workload (rank 0, N=1 only)
        cpu = None
        if world == 1:
            n_s = 1_000_000
            ys = x[:n_s].cpu().numpy()
            v, kind, cores = cpu_label_generation(ys, centers.cpu().numpy(), repeats=2)
            cpu = {"value": v, "unit": "rotations/s", "cores": cores, "kind": kind,
                   "sample": "first %d of the 10M rotations, K=%d, best of 2" % (n_s, K_DICT)}
        line = {
            "metric": METRIC, "value": value, "unit": "rotations/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (f32 screen, f64 exact re-check)",
            "data": "synthetic",
            "config": {"workload": "configs[1]: binDeltaGenerators label generation, 10M rotations per GPU, "
                                   "K=1000 dictionary, nearest key + residual delta",
                       "n_rotations_per_gpu": N_ROT, "K": K_DICT, "x_dtype": "f32", "label_dtype": "int64",
                       "l2": "inputs+outputs 320 MB per step > 126 MB L2 (no explicit flush)",
                       "algorithm": "key grid (candidate pruning), rebuilt every step"},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "rotations/s", "ms_per_step": float(ms_e2e),
                    "h2d_bytes_per_step": N_ROT * 12, "d2h_bytes_per_step": N_ROT * 20},
            "gpu_launches": 4 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": 277.4e6, "peak_source": peak_src,
                         "kernel": "assign_grid_kernel<float,3,false>", "kernel_ms": kernel_s * 1e3,
                         "algorithmic_bytes_per_launch": N_ROT * BYTES_PER_ROT,
                         "note": "32 B/rotation x 10M rotations per launch; traffic = dram read+write of "
                                 "one launch from profiles/r1d_ncu_assign.csv; the step also runs the 3 "
                                 "key-grid build launches (~50 us)" +
                                 ("; brute-force scan of the same step: %.2f ms" % brute_ms if brute_ms else "")},
            "cpu_baseline": cpu,
            "extras": ex,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
