#!/bin/bash
# ncu capture of the query kernel (assign f32 d=3, then Lloyd): raw metrics + source-level stalls
set -u
OUT=gpurun_out
TAG=${1:-q}
cat > /tmp/pq.py <<'P'
import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [os.path.join(ROOT, "multi-modal-regression_b200"), ROOT]
import torch
from bench import synth_rotations
from bdpose import ops, _lib as L
dev = torch.device("cuda", 0)
x = synth_rotations(10_000_000, 1000, dev)
c = synth_rotations(1000, 7, dev).double().contiguous()
g = ops.KeyGrid(c)
which = sys.argv[1]
if which == "assign":
    for _ in range(3):
        ops.assign_nearest(x, c, grid=g)
else:
    xd = x.double().contiguous(); lib = L.lib()
    lab = torch.full((10_000_000,), -1, dtype=torch.int32, device=dev)
    acc = torch.zeros(7002, dtype=torch.int64, device=dev)
    for _ in range(3):
        lib.bdp_kmeans_lloyd_step_grid(xd.data_ptr(), 10_000_000, 3, c.data_ptr(), 1000, g.buf.data_ptr(), g.nbytes,
                                       lab.data_ptr(), acc.data_ptr(), 29, acc[7000:].data_ptr(), None, 1, L.stream_ptr())
torch.cuda.synchronize()
P
for which in ${2:-assign lloyd}; do
  ncu --set full --clock-control none --import-source on -k regex:"query_kernel|assign_grid_kernel" -s 2 -c 1 -f -o /tmp/prof_$which python /tmp/pq.py $which > $OUT/ncu_${TAG}_$which.log 2>&1
  ncu -i /tmp/prof_$which.ncu-rep --page raw --csv > /tmp/raw_$which.csv 2>/dev/null
  ncu -i /tmp/prof_$which.ncu-rep --page source --csv > /tmp/src_$which.csv 2>/dev/null
  python profiles/top_stalls.py /tmp/src_$which.csv 45 > $OUT/stalls_${TAG}_$which.txt 2>&1; gzip -c /tmp/src_$which.csv > $OUT/src_${TAG}_$which.csv.gz
  python - <<P2 > $OUT/raw_${TAG}_$which.txt
import csv
rows=list(csv.reader(open('/tmp/raw_$which.csv')))
h=rows[0]; u=rows[1]; v=rows[2]
keep=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__registers_per_thread','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_op_read.sum','lts__t_sector_hit_rate.pct','smsp__average_warp_latency_per_inst_issued.ratio','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct','smsp__warp_issue_stalled_no_instruction_per_warp_active.pct','smsp__warp_issue_stalled_not_selected_per_warp_active.pct','smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct','smsp__warp_issue_stalled_membar_per_warp_active.pct','smsp__warp_issue_stalled_sleeping_per_warp_active.pct']
for k in keep:
    if k in h:
        i=h.index(k); print(k, v[i], u[i])
P2
done
