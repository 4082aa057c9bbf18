#!/bin/bash
# ncu capture of ONE steady-state E+M kernel of the k-means loop as the bench runs it (sorted rows,
# fixed-geometry grid, incremental M-step): launch index $2 (default 10) of query_kernel.
set -u
OUT=gpurun_out
TAG=${1:-loop}
SKIP=${2:-10}
cat > /tmp/pl.py <<'P'
import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [os.path.join(ROOT, "multi-modal-regression_b200"), ROOT]
import torch
from bench import kmeans_chunks, synth_rotations, N_ROT, N_CHUNKS, K_DICT
from bdpose import ops, kmeans
dev = torch.device("cuda", 0)
xs = kmeans_chunks(range(N_CHUNKS), dev)
init = synth_rotations(N_ROT // N_CHUNKS, 100, dev, torch.float64)[:K_DICT].clone()
fs = kmeans.FitSetup(xs, init, group=kmeans.LOCAL)
labels = torch.full((xs.shape[0],), -1, dtype=torch.int32, device=dev)
loop = kmeans.LloydLoop(fs.x, fs.centers, labels, fs.hb, ops.KeyGrid(fs.centers), kmeans.LOCAL, fs.tol_abs, box=fs.max_abs)
loop.launch(0, 14, False)
torch.cuda.synchronize()
P
timeout 120 python /tmp/pl.py > $OUT/plain_${TAG}.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"query_kernel" -s $SKIP -c 1 -f -o /tmp/prof_$TAG python /tmp/pl.py > $OUT/ncu_${TAG}.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > /tmp/raw_$TAG.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > /tmp/src_$TAG.csv 2>/dev/null
python profiles/top_stalls.py /tmp/src_$TAG.csv 40 > $OUT/stalls_${TAG}.txt 2>&1; gzip -c /tmp/src_$TAG.csv > $OUT/src_${TAG}.csv.gz
python - <<P2 > $OUT/raw_${TAG}.txt
import csv
rows=list(csv.reader(open('/tmp/raw_$TAG.csv')))
h=rows[0]; u=rows[1]; v=rows[2]
keep=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__registers_per_thread','launch__block_size','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_op_read.sum','lts__t_sector_hit_rate.pct','smsp__average_warp_latency_per_inst_issued.ratio','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct','smsp__warp_issue_stalled_no_instruction_per_warp_active.pct','smsp__warp_issue_stalled_not_selected_per_warp_active.pct','smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct','smsp__warp_issue_stalled_membar_per_warp_active.pct','smsp__warp_issue_stalled_sleeping_per_warp_active.pct']
for k in keep:
    if k in h:
        i=h.index(k); print(k, v[i], u[i])
P2
cat $OUT/raw_${TAG}.txt
