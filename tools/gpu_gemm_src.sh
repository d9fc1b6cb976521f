#!/bin/bash
# ncu source-level capture of ONE launch of the forward edge GEMM (M = E, N = 1792, K = 256): per-line stall samples.
out=gpurun_out; mkdir -p $out
python tools/gpu_gemm_one.py 301491 1792 256 1 6
ncu --set full --import-source on --clock-control none -k regex:gemm16_kernel --launch-skip 3 --launch-count 1 -f -o /tmp/gemm_src \
  python tools/gpu_gemm_one.py 301491 1792 256 1 3 > $out/ncu_gemm_src.log 2>&1
ncu -i /tmp/gemm_src.ncu-rep --page source --csv --print-source sass > $out/gemm_src_sass.csv 2>/dev/null
ncu -i /tmp/gemm_src.ncu-rep --page source --csv --print-source cuda > $out/gemm_src_cuda.csv 2>/dev/null
ncu -i /tmp/gemm_src.ncu-rep --page raw --csv > $out/gemm_src_raw.csv 2>/dev/null
ls -la $out/gemm_src*; gzip -f $out/gemm_src_sass.csv
