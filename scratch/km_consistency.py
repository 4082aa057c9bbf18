"""Hashes of the centres after 20 fixed iterations for the kernel variants (must all agree)."""
import os, sys, hashlib, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path[:0] = [os.path.join(ROOT, "multi-modal-regression_b200"), ROOT]
    import torch
    from bench import kmeans_chunks, synth_rotations, N_ROT, N_CHUNKS, K_DICT
    from bdpose import kmeans
    n = int(sys.argv[2])
    dev = torch.device("cuda", 0)
    xs = kmeans_chunks(range(N_CHUNKS), dev)[:n].contiguous()
    init = synth_rotations(N_ROT // N_CHUNKS, 100, dev, torch.float64)[:K_DICT].clone()
    for rep in range(2):
        r = kmeans.kmeans_lloyd(xs, init, fixed_iters=20, group=kmeans.LOCAL,
                                use_grid=os.environ.get("KM_GRID", "auto") != "0" and "auto")
        h = hashlib.sha256(r["centers"].cpu().numpy().tobytes()).hexdigest()[:12]
        hl = hashlib.sha256(r["labels"].cpu().numpy().tobytes()).hexdigest()[:12]
        print("  rep %d centers %s labels %s inertia %.9f" % (rep, h, hl, r["inertia"]))
    sys.exit(0)
n = sys.argv[1] if len(sys.argv) > 1 else "1250000"
for name, env in (("incremental", {}), ("full", {"BDPOSE_KMEANS_INCREMENTAL": "0"}),
                  ("brute force", {"KM_GRID": "0"})):
    print(name, flush=True)
    subprocess.run([sys.executable, __file__, "child", n], env=dict(os.environ, **env))
