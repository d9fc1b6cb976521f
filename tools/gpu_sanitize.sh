#!/bin/bash
# compute-sanitizer pass over the CUDA path (SURVEY.md §5 "race detection"): memcheck, racecheck, synccheck, initcheck
# on the golden cases (cfg1, yaml_l2, l3_trunc, norms, two-layer gamma_t), training-mode dropout, the C=256 model,
# the energy/force head, the equivariant heads and the GEMM layout tests.  Run on the GPU box:
#   gpurun --timeout 3000 -- bash tools/gpu_sanitize.sh
# Logs land in gpurun_out/sanitizer_<tool>.txt (copy the summaries to profiles/).
set -u
mkdir -p gpurun_out
SEL='(golden_forward_backward and (cfg1 or yaml_l2 or l3_trunc or norms_l3 or eu_mlp-)) or attention_dropout or cfg2_width or (head_energy_forces_golden and l2) or dipole_and_spatial or (gemm_layouts and (257 or 1000)) or fp16_split'
TOOLS=${1:-"memcheck racecheck synccheck initcheck"}
for tool in $TOOLS; do
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check no"
  start=$(date +%s)
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 40 --error-exitcode 9 \
    --log-file gpurun_out/sanitizer_${tool}_raw.txt \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_${tool}_pytest.txt 2>&1
  rc=$?
  end=$(date +%s)
  {
    echo "# compute-sanitizer --tool $tool  (rc=$rc, $((end-start)) s)  selection: $SEL"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|Barrier error|========= (Error|Warning)" gpurun_out/sanitizer_${tool}_raw.txt | sort | uniq -c | sort -rn | head -40
    echo "# pytest tail:"
    tail -4 gpurun_out/sanitizer_${tool}_pytest.txt
  } > gpurun_out/sanitizer_${tool}.txt
  head -c 20000 gpurun_out/sanitizer_${tool}_raw.txt > gpurun_out/sanitizer_${tool}_head.txt
  cat gpurun_out/sanitizer_${tool}.txt
done
