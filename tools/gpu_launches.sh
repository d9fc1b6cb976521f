#!/bin/bash
# per-launch time list of one bench step under the current environment:  tools/gpu_launches.sh <tag> [topN]
tag=$1; top=${2:-12}
out=gpurun_out; mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1140 -c 400 --csv --log-file $out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > $out/ncu_launch_$tag.log 2>&1
echo "== $tag"; python tools/agg_launches.py $out/launches_$tag.csv $top
