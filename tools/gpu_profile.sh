#!/bin/bash
# Round profiling pass on the GPU box (run through gpurun):  tools/gpu_profile.sh <tag>
# 1. bench line, 2. launch list of one step (gpu__time_duration), 3. ncu --set full of one step's GEMM launches and of
# one step's graph kernels (gata / htr).  The .ncu-rep files are exported to CSV (raw page) on the box: gpurun_out/ is
# capped at 64 MiB.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
python bench.py --steps 10 --warmup 3 --no-workloads > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
python tools/bench_table.py $out/bench_$tag.json | head -12
L=$(python -c "import json;print(json.load(open('$out/bench_$tag.json'))['gpu_launches'])")
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((3 * L + 60)) -c $L --csv --log-file $out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > $out/ncu_launch_$tag.log 2>&1
python tools/agg_launches.py $out/launches_$tag.csv 40 > $out/launches_${tag}_summary.txt; head -30 $out/launches_${tag}_summary.txt
ncu --set full --clock-control none -k regex:'gata_|htr_' --launch-skip 29 --launch-count 13 -f -o /tmp/full_graph_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > $out/ncu_graph_$tag.log 2>&1
ncu -i /tmp/full_graph_$tag.ncu-rep --page raw --csv > $out/full_graph_$tag.csv 2>/dev/null
python tools/ncu_summary.py $out/full_graph_$tag.csv > $out/full_graph_${tag}_summary.txt; cat $out/full_graph_${tag}_summary.txt
ncu --set full --clock-control none -k regex:'gemm16' --launch-skip 110 --launch-count 110 -f -o /tmp/full_gemm_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > $out/ncu_gemm_$tag.log 2>&1
ncu -i /tmp/full_gemm_$tag.ncu-rep --page raw --csv > $out/full_gemm_$tag.csv 2>/dev/null
python tools/ncu_summary.py $out/full_gemm_$tag.csv > $out/full_gemm_${tag}_summary.txt; head -40 $out/full_gemm_${tag}_summary.txt
gzip -f $out/full_gemm_$tag.csv $out/full_graph_$tag.csv
ls -la $out | tail -12; du -sh $out
