#!/bin/bash
# Run on the GPU box (gpurun): launch list of the bench command + one `ncu --set full` capture per hot
# kernel, exported to small CSVs (the .ncu-rep files stay on the box: gpurun_out/ is capped at 64 MiB).
# usage: bash profiles/capture.sh <tag> [families: "assign loss lloyd head" (default all); "nolist" skips step 1]
set -u
TAG=${1:-r2}
FAMILIES=${2:-"list assign loss lloyd head"}
OUT=gpurun_out
METRICS="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"
# 1) launch list of the benchmark command (after it exited 0 without ncu)
if [[ " $FAMILIES " == *" list "* ]]; then
timeout 200 python bench.py --steps 5 --warmup 3 --no-extras > $OUT/plain_bench_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_bench_$TAG.csv timeout 400 python bench.py --steps 5 --warmup 3 --no-extras > $OUT/ncu_bench_$TAG.log 2>&1
fi
# 2) full-set captures, a few launches per kernel family
python profiles/prof_targets.py assign loss lloyd head > $OUT/plain_prof_$TAG.log 2>&1 || exit 1
for spec in "assign:query_kernel|keygrid:4" "loss:bd_loss_kernel:2" "lloyd:query_kernel|keygrid|kmeans_xfin|cell_keys|gather_rows|RadixSort|scatter_i32|fit_stats:60" "head:gemm_tf32_kernel|bn_relu|fc3_:38"; do
  which=${spec%%:*}; rest=${spec#*:}; rx=${rest%%:*}; cnt=${rest##*:}
  [[ " $FAMILIES " == *" $which "* ]] || continue
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -f -o /tmp/prof_${which} \
      python profiles/prof_targets.py $which > $OUT/ncu_prof_${which}_$TAG.log 2>&1
  ncu -i /tmp/prof_${which}.ncu-rep --page raw --csv > /tmp/raw_${which}.csv 2>/dev/null
  python profiles/pick_metrics.py /tmp/raw_${which}.csv "$METRICS" > $OUT/ncu_${which}_$TAG.csv
done
ls -la $OUT | tail -20
