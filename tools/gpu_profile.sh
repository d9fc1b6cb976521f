#!/bin/bash
# Round profiling pass on the GPU box (run through gpurun):  tools/gpu_profile.sh <tag> [notest]
# 1. parity tests, 2. bench line, 3. launch list (gpu__time_duration), 4. ncu --set full of the
# graph kernels (one whole step's worth of gata/htr launches) and of a window of GEMM launches.
# The .ncu-rep files are exported to CSV (raw page) on the box and removed: gpurun_out/ is capped at 64 MiB.
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
if [ "$2" != "notest" ]; then
  python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
fi
python bench.py --steps 10 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
tail -c 600 $out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1140 -c 380 --csv --log-file $out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none -k regex:'gata_|htr_' --launch-skip 63 --launch-count 13 -f -o /tmp/full_graph_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_graph_$tag.log 2>&1
ncu -i /tmp/full_graph_$tag.ncu-rep --page raw --csv > $out/full_graph_$tag.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'gemm3x' --launch-skip 345 --launch-count 30 -f -o /tmp/full_gemm_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_gemm_$tag.log 2>&1
ncu -i /tmp/full_gemm_$tag.ncu-rep --page raw --csv > $out/full_gemm_$tag.csv 2>/dev/null
ls -la $out | tail -12; du -sh $out
