#!/bin/bash
# ncu --set full + source page of ONE key-grid cell kernel (and one exchange kernel) of the k-means
# loop exactly as the bench runs it (fixed-geometry grid, occupied cells only): launch $2 of the run.
# usage (on the GPU box): bash scratch/prof_cell.sh <tag> [skip]
tag=${1:-cell}; skip=${2:-10}
out=gpurun_out
cmd="python bench.py --steps 5 --warmup 3 --no-extras"
timeout 120 $cmd > $out/plain_$tag.log 2>&1 || { echo "plain run failed"; exit 1; }
for k in keygrid_cell_kernel kmeans_xfin_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip --launch-count 1 \
      -f -o /tmp/$tag.$k $cmd > $out/ncu_$tag.$k.log 2>&1
  ncu -i /tmp/$tag.$k.ncu-rep --page raw > $out/raw_$tag.$k.txt 2>/dev/null
  ncu -i /tmp/$tag.$k.ncu-rep --page source --csv 2>/dev/null | gzip > $out/src_$tag.$k.csv.gz
done
ls -la $out | grep $tag
