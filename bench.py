#!/usr/bin/env python
"""bench.py — headline benchmark of the bin-and-delta pose hot path on B200.

Workload (BASELINE.json configs[2], the configuration the `metric` is quoted on): k-means
pose-dictionary learning (learnKmeansDictionary.py:41-42) on 10 M random rotations (axis-angle, fp64 as
the reference feeds scikit-learn), K = 1000, from an explicit init (the first K rotations).  One STEP =
one Lloyd iteration (E-step + M-step + exchange) over all 10 M rotations; `--steps K` iterations are
timed from the init after `--warmup W` untimed ones.  With --gpus N the SAME 10 M rotations are sharded
over the ranks (strong scaling); the per-iteration exchange of the cluster sums is fused into the
finalise kernel over symmetric memory (no NCCL call in the loop).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extras] [--multi-extras]

Prints ONE JSON line (rank 0).  `value` = rotation-iterations/s with the shard resident in HBM;
`e2e` = the same metric through the public API (bdpose.kmeans.KMeans.fit on a pinned HOST array: H2D of
the shard, centring, the iterations, final E-step, D2H of labels and centres inside the timed region);
`roofline` = the E+M kernel against the measured HBM peak; `cpu_baseline` = scikit-learn's KMeans.fit
(the reference's own call, same explicit init) on the host cores, bounded sample; `parity` = the
centres of the timed fit hashed on every rank (ranks agree; equal to a single-GPU fit of the same
data).  `extras` carries the other BASELINE.json configs (label generation, fused loss, evaluation,
head + loss) with their own roofline / CPU numbers; they run at N = 1 (at N > 1 only with --multi-extras).
"""
import argparse
import csv
import hashlib
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
N_CHUNKS = 8                         # the data set is 8 seeded chunks: any of 1/2/4/8 ranks owns whole chunks
BYTES_PER_ROT_ITER = 24 + 4          # fp64 rotation in, int32 label out
BYTES_PER_ROT_LABEL = 12 + 8 + 12    # label generation: fp32 y in, int64 bin out, fp32 residual out
METRIC = "k-means rotation-iterations/s (Lloyd E+M over 10M rotations, K=1000, sharded at 1/2/4/8 GPUs)"
WORKLOAD = ("configs[2]: k-means pose-dictionary learning, 10M random rotations total (fp64), K=1000, "
            "explicit init = first K rotations, fixed Lloyd iterations, strong scaling")


def bench_config(steps):
    """The workload both arms (this one and --impl reference) run: identical on both lines."""
    return {"workload": WORKLOAD, "n_rotations_total": N_ROT, "K": K_DICT, "x_dtype": "f64",
            "init": "first K rotations (explicit)", "iterations_timed": steps,
            "l2": "inputs larger than L2: the 10M fp64 rotations + labels are 280 MB per pass (126 MB L2); "
                  "sharded over 4-8 GPUs a shard fits the L2 and stays there across iterations as in any "
                  "real fit - ms_per_step_l2_flushed re-times the iterations with a 256 MB write in between"}


def synth_rotations(n, seed, device, dtype=torch.float32):
    """Uniform rotations on SO(3) as axis-angle [n,3]."""
    g = torch.Generator(device=device).manual_seed(seed)
    q = torch.randn(n, 4, device=device, generator=g, dtype=dtype)
    q = q / q.norm(dim=1, keepdim=True)
    w = q[:, :1].abs().clamp(max=1)
    v = q[:, 1:] * torch.sign(q[:, :1])
    return (v / v.norm(dim=1, keepdim=True).clamp_min(1e-30) * (2 * torch.acos(w))).contiguous()


def kmeans_chunks(chunk_ids, device):
    """The fp64 rotations of the given chunks of the 10 M-rotation data set (chunk c = seed 100 + c,
    always generated with the same launch shape, so the values do not depend on the world size)."""
    per = N_ROT // N_CHUNKS
    return torch.cat([synth_rotations(per, 100 + c, device, torch.float64) for c in chunk_ids])


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


def profile_traffic(fname, kernel_substr):
    """dram read + write bytes per launch of a kernel from a committed ncu export under profiles/
    (columns picked by profiles/pick_metrics.py), or None."""
    path = os.path.join(ROOT, "profiles", fname)
    try:
        if fname.endswith(".txt"):
            # "metric value unit" lines of ONE kernel launch (scratch/prof_loop.sh: a steady-state E+M
            # kernel of the loop as the bench runs it)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = 0.0
            for line in open(path):
                f = line.split()
                if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(f[1]) * scale.get(f[2], 1.0)
            return tot or None
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = [float(r[ir]) * scale.get(units[ir], 1.0) + float(r[iw]) * scale.get(units[iw], 1.0)
                for r in rows[2:] if kernel_substr in r[0]]
        return (sum(vals) / len(vals)) if vals else None
    except Exception:
        return None


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


# ---------------------------------------------------------------------------------------------------
# CPU legs: the reference's own calls on the host cores
# ---------------------------------------------------------------------------------------------------
def sklearn_fit_rate(X, init, iters_a, iters_b):
    """rotation-iterations/s of scikit-learn's KMeans.fit (learnKmeansDictionary.py:41-42 with the
    explicit init the parity runs use, n_init=1, Lloyd) as the MARGINAL cost of an iteration: the
    difference of two fits with iters_a < iters_b iterations (tol=0: no early stop), which cancels
    the one-off work (centring, norms, the final E-step).  Returns (rate, seconds per iteration)."""
    from sklearn.cluster import KMeans

    def fit(n):
        t0 = time.perf_counter()
        km = KMeans(n_clusters=init.shape[0], init=init, n_init=1, max_iter=n, tol=0.0,
                    algorithm="lloyd").fit(X)
        return time.perf_counter() - t0, km.n_iter_
    ta, na = fit(iters_a)
    tb, nb = fit(iters_b)
    per = (tb - ta) / max(nb - na, 1)
    if per <= 0:
        per = tb / max(nb, 1)
    return X.shape[0] / per, per


def cpu_label_generation(y, centers):
    """sklearn predict + numpy residual (binDeltaGenerators.py:27-31).  Returns rotations/s."""
    from sklearn.cluster import KMeans
    km = KMeans(n_clusters=centers.shape[0], init=centers, n_init=1, max_iter=1)
    km.cluster_centers_ = np.ascontiguousarray(centers)
    km._n_threads = os.cpu_count() or 1
    km.n_features_in_ = centers.shape[1]
    km._n_init = 1

    def run():
        b = km.predict(y.astype(np.float64))
        return b, (y - centers[b, :]).astype(np.float32)
    run()
    t0 = time.perf_counter()
    run()
    return y.shape[0] / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    """--impl reference: scikit-learn's KMeans.fit on the box's host cores — the SAME 10 M rotations,
    K and explicit init as the GPU arm; a step = one Lloyd iteration.  Rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    X = kmeans_chunks(range(N_CHUNKS), "cpu").numpy()
    n_env = int(os.environ.get("BDP_BENCH_REF_ROTATIONS", "0"))     # CPU test suite: a smaller sample
    if 0 < n_env < X.shape[0]:
        X = X[:n_env]
    init = X[:K_DICT].copy()
    w = max(args.warmup, 1)
    # two fits: W iterations (also the warm-up of the BLAS/OpenMP pools) and W + K iterations
    t0 = time.perf_counter()
    rate, per = sklearn_fit_rate(X, init, w, w + args.steps)
    wall = time.perf_counter() - t0
    sample = ("%s %d rotations, K=%d: KMeans(init=first K rows, n_init=1, tol=0, lloyd).fit with %d and "
              "%d iterations, marginal cost of the %d extra iterations (%.0f s of CPU wall in total)"
              % ("all" if X.shape[0] == N_ROT else "first", X.shape[0], K_DICT, w, w + args.steps,
                 args.steps, wall))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "rotation-iterations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": bench_config(args.steps),
            "cpu_baseline": {"value": rate, "unit": "rotation-iterations/s", "cores": cores,
                             "kind": "reference", "sample": sample},
            "e2e": {"value": rate, "unit": "rotation-iterations/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def kmeans_k200_leg(dev, K=200, steps=20):
    """The headline loop at K = 200 (one GPU): ms per Lloyd iteration, CUDA events around the loop."""
    from bdpose import ops, kmeans
    xs = kmeans_chunks(range(N_CHUNKS), dev)
    init = synth_rotations(N_ROT // N_CHUNKS, 100, dev, torch.float64)[:K].clone()
    fs = kmeans.FitSetup(xs, init, group=kmeans.LOCAL)
    labels = torch.full((xs.shape[0],), -1, dtype=torch.int32, device=dev)
    loop = kmeans.LloydLoop(fs.x, fs.centers, labels, fs.hb, ops.KeyGrid(fs.centers, build=False), kmeans.LOCAL,
                            fs.tol_abs, box=fs.max_abs)
    ms = 0.0
    for rep in range(2):
        loop.reset(fs.centers)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop.launch(0, steps, False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    return {"ms_per_iteration": ms, "rotation_iterations_per_s": N_ROT / (ms * 1e-3), "K": K,
            "iterations": steps, "hbm_frac_whole_iteration": N_ROT * BYTES_PER_ROT_ITER / (ms * 1e-3) / 1e9
            / measured_peaks()[0]["hbm_gbs"], "occupied_coarse_cells": [loop.n_cells, getattr(loop, "n_coarse", 0)]}


# ---------------------------------------------------------------------------------------------------
# extras: the other BASELINE.json configs, short runs
# ---------------------------------------------------------------------------------------------------
def extras(dev, peaks, rank, world, args):
    from bdpose import ops, _lib as L
    import binDeltaGenerators as G
    import torch.distributed as dist
    out = {}
    hbm = peaks["hbm_gbs"]
    # ---- config 4: data-parallel head step (all ranks) ------------------------------------------------
    if world > 1:
        import bench_head
        out.update(bench_head.bench_dp(dev, world))
    # ---- config 1 (label generation, every rank its own 10 M rotations: weak scaling) ---------------
    x = synth_rotations(N_ROT, 1000 + rank, dev)
    centers = synth_rotations(K_DICT, 7, dev).double().contiguous()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([timed(lambda: ops.assign_nearest(x, centers, want_residual=True), 20, 3)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    grid = ops.KeyGrid(centers)
    ms_q = timed(lambda: ops.assign_nearest(x, centers, grid=grid), 20, 3)
    by = N_ROT * BYTES_PER_ROT_LABEL
    out["label_generation_10M_K1000"] = {
        "rotations_per_s": N_ROT * world / (float(ms) * 1e-3), "ms_per_step": float(ms), "scaling": "weak",
        "roofline": {"bound": "hbm", "kernel": "assign query kernel", "kernel_ms": ms_q,
                     "achieved": by / (ms_q * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": by / (ms_q * 1e-3) / 1e9 / hbm, "algorithmic_bytes_per_launch": by,
                     "traffic": profile_traffic("r2f_ncu_assign.csv", "query_kernel")}}
    if rank != 0:
        return out
    if world == 1:
        # host-to-host label generation (pinned buffers, 3-stream chunk pipeline) + the CPU call
        y_host = x.cpu().pin_memory()
        bin_host = torch.empty(N_ROT, dtype=torch.int64).pin_memory()
        res_host = torch.empty(N_ROT, 3, dtype=torch.float32).pin_memory()
        ms_h = timed(lambda: G.assign_labels_host(y_host, centers, out_bin=bin_host, out_res=res_host), 5, 2)
        out["label_generation_10M_K1000"]["e2e_host_to_host"] = {
            "rotations_per_s": N_ROT / (ms_h * 1e-3), "ms": ms_h, "h2d_bytes": N_ROT * 12, "d2h_bytes": N_ROT * 20}
        del y_host, bin_host, res_host
        try:
            v = cpu_label_generation(x[:1_000_000].cpu().numpy(), centers.cpu().numpy())
            out["label_generation_10M_K1000"]["cpu_baseline"] = {
                "value": v, "unit": "rotations/s", "cores": os.cpu_count(), "kind": "reference",
                "sample": "sklearn predict + numpy residual on the first 1M rotations"}
        except Exception as e:      # pragma: no cover
            out["label_generation_10M_K1000"]["cpu_baseline"] = {"error": str(e)}
    # clustered dictionary (Pascal-like azimuth / elevation / tilt poses): list lengths + slow path
    out["label_generation_clustered"] = clustered_leg(dev, ops, hbm)
    if world == 1:
        # configs[2] also names K = 200: the same fixed-work fit at the smaller dictionary
        out["kmeans_10M_K200"] = kmeans_k200_leg(dev)
    # ---- config 1b: fused loss fwd+bwd, 1 M rows, K=200, through autograd ---------------------------
    B, K = 1_000_000, 200
    score = torch.randn(B, K, device=dev, requires_grad=True)
    bins = torch.randint(0, K, (B,), device=dev)
    delta = (torch.randn(B, 3, device=dev) * 0.2).requires_grad_(True)
    target = x[:B].contiguous()
    keys = centers[:K].float().contiguous()

    def loss_step():
        score.grad = None; delta.grad = None
        lc, lr, _ = ops.bd_loss(score, bins, delta, target, keys, L.POSE_GEODESIC_AA, True)
        (lc + 0.7 * lr).backward()
    ms_a = timed(loss_step, 10, 3)
    ms_r = timed(lambda: ops.bd_loss_raw(score.detach(), bins, delta.detach(), target, keys,
                                         L.POSE_GEODESIC_AA, True), 10, 3)
    by = B * (2 * K * 4 + 8 + 12 + 12 + 12 + 8)
    out["fused_loss_1M_K200"] = {
        "samples_per_s": B / (ms_a * 1e-3), "ms_autograd_fwd_bwd": ms_a, "ms_kernel": ms_r, "bytes": by,
        "roofline": {"bound": "hbm", "kernel": "bd_loss_kernel", "achieved": by / (ms_r * 1e-3) / 1e9,
                     "peak": hbm, "unit": "GB/s", "frac": by / (ms_r * 1e-3) / 1e9 / hbm,
                     "frac_through_autograd": by / (ms_a * 1e-3) / 1e9 / hbm,
                     "traffic": profile_traffic("r2f_ncu_loss.csv", "bd_loss")}}
    del score, delta
    # ---- config 5: evaluation over 1 M predictions -------------------------------------------------
    a = x[:1_000_000].contiguous(); b = x[1_000_000:2_000_000].contiguous()
    labels = torch.randint(0, 12, (1_000_000,), device=dev)

    def ev():
        e = ops.geodesic_error_deg(a, b)
        ops.error_stats(e, labels, 12)
    ms = timed(ev, 10, 3)
    out["eval_1M"] = {"pairs_per_s": 1e6 / (ms * 1e-3), "ms": ms}
    ms = timed(lambda: ops.geodesic_error_deg(a, b), 10, 3)
    out["eval_1M"]["error_kernel_ms"] = ms
    out["eval_1M"]["error_kernel_hbm_frac"] = 1_000_000 * (12 + 12 + 8) / (ms * 1e-3) / 1e9 / hbm
    del x
    # ---- config 1 / 4: head + loss fwd/bwd -------------------------------------------------------------
    import bench_head
    out.update(bench_head.bench(dev, peaks))
    return out


def clustered_leg(dev, ops, hbm):
    """Label generation against a CLUSTERED dictionary: Pascal3D+-like poses (azimuth anywhere,
    elevation and camera tilt concentrated near 0) -> k-means dictionary -> key-grid statistics."""
    from bdpose import kmeans
    g = torch.Generator(device=dev).manual_seed(5)
    n = 2_000_000
    eul = torch.stack([torch.rand(n, device=dev, generator=g, dtype=torch.float64) * 360.0,
                       torch.randn(n, device=dev, generator=g, dtype=torch.float64) * 12.0 + 5.0,
                       torch.randn(n, device=dev, generator=g, dtype=torch.float64) * 8.0], 1)
    y, _ = ops.euler_to_pose(eul)
    r = kmeans.kmeans_lloyd(y, y[:200].clone(), max_iter=30, group=kmeans.LOCAL)
    centers = r["centers"]
    yf = y.float().contiguous()
    ms = timed(lambda: ops.assign_nearest(yf, centers, want_residual=True), 10, 3)
    stats = ops.keygrid_stats(ops.KeyGrid(centers), yf) if hasattr(ops, "keygrid_stats") else None
    return {"n": n, "K": 200, "kmeans_iterations": r["n_iter"], "rotations_per_s": n / (ms * 1e-3),
            "ms": ms, "hbm_frac": n * BYTES_PER_ROT_LABEL / (ms * 1e-3) / 1e9 / hbm, "key_grid": stats}


def sha(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--multi-extras", action="store_true",
                    help="with --gpus N > 1: also run the extras' multi-rank legs")
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
    from bdpose import ops, kmeans, _lib as L
    lib = L.lib()
    peaks, peak_src = measured_peaks()
    steps, warmup = args.steps, max(args.warmup, 3)
    if N_CHUNKS % world:
        raise SystemExit("bench.py: --gpus must divide %d" % N_CHUNKS)

    # this rank's shard of the 10 M rotations + the explicit init (first K rotations of chunk 0)
    per_rank = N_CHUNKS // world
    xs = kmeans_chunks(range(rank * per_rank, (rank + 1) * per_rank), dev)
    init = synth_rotations(N_ROT // N_CHUNKS, 100, dev, torch.float64)[:K_DICT].clone()
    n_local = xs.shape[0]
    fs = kmeans.FitSetup(xs, init)                     # centring, tolerance, fixed-point scale
    labels = torch.full((n_local,), -1, dtype=torch.int32, device=dev)
    grid = ops.KeyGrid(fs.centers)
    loop = kmeans.LloydLoop(fs.x, fs.centers, labels, fs.hb, grid, None, fs.tol_abs, box=fs.max_abs)
    n_cells, n_coarse = loop.n_cells, getattr(loop, "n_coarse", 0)

    def fit_region(n):
        loop.reset(fs.centers)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop.launch(0, n, False)                       # n iterations queued back to back, no host sync
        e1.record()
        torch.cuda.synchronize()
        loop.n_iter = n
        return e0.elapsed_time(e1)
    fit_region(warmup)
    with ClockSampler(local) as clk:
        ms_total = torch.tensor([fit_region(steps)], device=dev, dtype=torch.float64)
    st = loop.status()
    if st.state != L.KMEANS_RUNNING or st.iter_done != steps:
        raise SystemExit("bench.py: the timed fit stopped early (state %d after %d iterations)"
                         % (st.state, st.iter_done))
    centers_timed = loop.centers.clone()
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    clocks = clk.summary()
    # the timed region is a few ms: keep the same iterations running for ~0.7 s so that nvidia-smi
    # sees the clocks and throttle reasons under this load (the SAME number of repetitions on every
    # rank: each one contains collectives)
    reps = max(1, min(100, int(700.0 / max(float(ms_total), 1.0))))
    with ClockSampler(local) as clk2:
        for _ in range(reps):
            fit_region(steps)
    clk2.rows = clk.rows + clk2.rows
    clocks = clk2.summary()
    clocks["note"] = ("timed region is %.1f ms; sampled over it plus %d more repetitions of the same "
                      "iterations" % (float(ms_total), reps))
    ms_step = float(ms_total) / steps
    value = N_ROT / (ms_step * 1e-3)

    # ---- parity: every rank holds the same centres; they equal a single-GPU fit of all the data ----
    h = sha(centers_timed)
    parity = {"centers_sha256_16": h, "exchange": loop.mode}
    if world > 1:
        hs = [None] * world
        dist.all_gather_object(hs, h)
        parity["ranks_agree"] = len(set(hs)) == 1
        if rank == 0:
            xa = kmeans_chunks(range(N_CHUNKS), dev)
            fa = kmeans.FitSetup(xa, init, group=kmeans.LOCAL)
            la = torch.full((xa.shape[0],), -1, dtype=torch.int32, device=dev)
            single = kmeans.LloydLoop(fa.x, fa.centers, la, fa.hb, ops.KeyGrid(fa.centers), kmeans.LOCAL,
                                      fa.tol_abs, box=fa.max_abs)
            single.launch(0, steps, False)
            torch.cuda.synchronize()
            single.n_iter = steps
            parity["equals_single_gpu"] = sha(single.centers) == h
            del xa, fa, la, single
        dist.barrier()
    else:
        parity["ranks_agree"] = True
        parity["equals_single_gpu"] = True

    # ---- dominant kernel: the E+M kernel of every iteration of the SAME fit, bracketed by CUDA events
    #      on its stream inside bdp_kmeans_run (the first iteration accumulates every rotation, the later
    #      ones only move the rotations whose label changed) --------------------------------------------
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4 * steps)]
    for e in evs:
        e.record()                                       # creates the underlying cudaEvent_t
    loop.reset(fs.centers)
    torch.cuda.synchronize()
    loop.launch(0, steps, False, em_events=evs)
    torch.cuda.synchronize()
    em_ms = [evs[4 * i + 1].elapsed_time(evs[4 * i + 2]) for i in range(steps)]
    kernel_ms = sum(em_ms) / steps
    anatomy = {"grid_build_ms": sum(evs[4 * i].elapsed_time(evs[4 * i + 1]) for i in range(steps)) / steps,
               "e_m_kernel_ms": kernel_ms,
               "exchange_finalise_ms": sum(evs[4 * i + 2].elapsed_time(evs[4 * i + 3]) for i in range(steps)) / steps,
               "note": "mean per iteration on this rank, CUDA events inside bdp_kmeans_run; with several ranks "
                       "the E+M kernel first waits for every rank's grid slab and the exchange kernel for "
                       "every rank's sums, so rank skew shows up there"}
    grid.rebuild(fs.centers)
    build_ms = timed(lambda: grid.rebuild(), max(steps, 10), 3)
    achieved = n_local * BYTES_PER_ROT_ITER / (kernel_ms * 1e-3) / 1e9

    # ---- iterations with an L2 flush in between (reported next to the back-to-back number) ----------
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    loop.reset(fs.centers)
    loop.launch(0, 1, False)
    acc_ms = 0.0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        flush.fill_(i & 1)
        ev[i][0].record()
        loop.launch(1 + i, 1, False)
        ev[i][1].record()
    torch.cuda.synchronize()
    acc_ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    fl = torch.tensor([acc_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(fl, op=dist.ReduceOp.MAX)
    ms_step_flushed = float(fl)
    del flush

    # ---- e2e: public API on a pinned HOST array (this rank's shard), labels + centres back on the host
    x_host = xs.cpu().pin_memory()
    init_np = init.cpu().numpy()
    grp = dist.group.WORLD if world > 1 else None

    def e2e_fit():
        km = kmeans.KMeans(n_clusters=K_DICT, init=init_np, n_init=1, max_iter=steps, fixed_iters=steps,
                           group=grp, device=dev)
        km.copy_labels = False           # labels_ stays in the pinned host buffer the D2H copy filled
        km.fit(x_host)
        return km
    e2e_fit()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    k_e2e = 3
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k_e2e):
        km = e2e_fit()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = torch.tensor([e0.elapsed_time(e1) / k_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_val = N_ROT * steps / (float(ms_e2e) * 1e-3)
    e2e_labels_ok = bool(km.labels_.shape[0] == n_local and km.cluster_centers_.shape == (K_DICT, 3))
    del x_host

    del loop, labels, xs, fs
    ex = {}
    # extras run on one GPU by default.  Under torchrun their multi-rank legs (data-parallel head step,
    # label generation on every rank) are opt-in (--multi-extras): the strong-scaling line the driver
    # records must not depend on them, and the driver truncates them out of its N>1 records anyway
    if not args.no_extras and (world == 1 or args.multi_extras):
        try:
            ex = extras(dev, peaks, rank, world, args)
        except Exception as e:
            if world > 1:
                raise           # the other ranks sit in a collective: fail loudly
            ex = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        cpu = None
        if world == 1:
            # CPU baseline on a bounded sample of the same workload (rank 0, N=1 only)
            n_s = 2_000_000
            Xs = kmeans_chunks([0, 1], "cpu").numpy()[:n_s]
            try:
                rate, per = sklearn_fit_rate(Xs, init_np, 2, 6)
                cpu = {"value": rate, "unit": "rotation-iterations/s", "cores": os.cpu_count() or 1,
                       "kind": "reference",
                       "sample": "scikit-learn KMeans.fit (explicit init, n_init=1, lloyd, tol=0) on the first "
                                 "%d of the 10M rotations, K=%d: marginal cost per iteration between fits of 2 "
                                 "and 6 iterations (%.2f s per iteration)" % (n_s, K_DICT, per)}
            except Exception as e:      # pragma: no cover
                cpu = {"error": str(e)}
        line = {
            "metric": METRIC, "value": value, "unit": "rotation-iterations/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 (f32 screen of the candidate keys, f64 exact re-check; int64 fixed-point sums)",
            "data": "synthetic",
            "config": bench_config(steps),
            "notes": {"n_rotations_per_gpu": n_local,
                      "l2": "shard per GPU = %d MB fp64 + %d MB labels; > 126 MB L2 at 1-2 GPUs, L2-resident "
                            "across iterations at 4-8 GPUs as in any real fit (no flush in the timed loop; "
                            "ms_per_step_l2_flushed re-times the iterations one by one with a 256 MB write "
                            "in between)" % (n_local * 24 // 1_000_000, n_local * 4 // 1_000_000),
                      "algorithm": "key grid (candidate pruning) over the bounding box of the rotations, its "
                                   "%d occupied coarse cells (of %d) rebuilt every iteration, sharded over "
                                   "the ranks; exact int64 fixed-point cluster sums, incremental M-step; "
                                   "exchange %s" % (n_cells, n_coarse, parity["exchange"])},
            "ms_per_step_l2_flushed": ms_step_flushed,
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "rotation-iterations/s", "ms_per_fit": float(ms_e2e),
                    "iterations_per_fit": steps, "h2d_bytes_per_step": n_local * 24 // steps,
                    "d2h_bytes_per_step": (n_local * 4 + K_DICT * 24) // steps,
                    "h2d_bytes_per_fit": n_local * 24, "d2h_bytes_per_fit": n_local * 4 + K_DICT * 24,
                    "outputs_ok": e2e_labels_ok,
                    "note": "KMeans.fit on a pinned host shard: H2D, centring / variance, the iterations, "
                            "final E-step + inertia, labels and centres D2H; a step is one iteration, so "
                            "the per-step byte counts are the per-fit counts divided by the iterations"},
            # per iteration: key-grid cell kernel, E+M kernel, exchange+finalise kernel; one header
            # kernel per bdp_kmeans_run call
            "gpu_launches": 3 * steps + 1,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"],
                         "traffic": profile_traffic("r2f_em_steady_raw.txt", "query_kernel"),
                         "peak_source": peak_src, "kernel": "Lloyd E+M kernel (assign query, fp64, accumulate)",
                         "kernel_ms": kernel_ms, "kernel_ms_first_iteration": em_ms[0],
                         "kernel_ms_last_iteration": em_ms[-1], "grid_build_ms_unsharded": build_ms,
                         "algorithmic_bytes_per_launch": n_local * BYTES_PER_ROT_ITER,
                         "note": "28 B/rotation-iteration x rotations of one rank per launch, duration = mean over "
                                 "the E+M kernels of the timed fit's iterations (CUDA events inside the loop); "
                                 "an iteration also runs the key-grid cell kernel (occupied cells only, sharded "
                                 "over the ranks) and the exchange+finalise kernel; grid_build_ms_unsharded is a "
                                 "full dictionary-box build (bdp_keygrid_build) for reference"},
            "iteration_anatomy": anatomy,
            "parity": parity,
            "cpu_baseline": cpu,
            "extras": ex,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
