#!/bin/bash
# quick GPU pass: parity tests + bench line + per-launch time list.  tools/gpu_quick.sh <tag> [pytest-args]
tag=${1:-q}
out=gpurun_out
mkdir -p $out
shift
timeout 900 python -m pytest tests -m gpu -x -q "$@" > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
cut -c1-400 $out/bench_$tag.json; tail -3 $out/bench_$tag.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1140 -c 400 --csv --log-file $out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_launch_$tag.log 2>&1
python tools/agg_launches.py $out/launches_$tag.csv 16
