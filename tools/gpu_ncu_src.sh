#!/bin/bash
# source-level ncu capture of one kernel:  tools/gpu_ncu_src.sh <tag> <kernel-regex> <skip> -- <cmd...>
tag=$1; rx=$2; skip=$3; shift 4
out=gpurun_out; mkdir -p $out
ncu --set full --import-source on --clock-control none -k regex:$rx --launch-skip $skip --launch-count 1 -f -o /tmp/src_$tag "$@" > $out/ncu_src_$tag.log 2>&1
ncu -i /tmp/src_$tag.ncu-rep --page raw --csv > $out/src_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/src_$tag.ncu-rep --page source --csv --print-source sass > $out/src_${tag}_sass.csv 2>/dev/null
ncu -i /tmp/src_$tag.ncu-rep --page source --csv --print-source cuda > $out/src_${tag}_cuda.csv 2>/dev/null
ls -la $out/src_${tag}_*
