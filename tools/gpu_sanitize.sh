#!/bin/bash
# compute-sanitizer pass over the CUDA path (SURVEY.md §5 "race detection"): memcheck, racecheck, synccheck, initcheck
# on the golden cases (cfg1, yaml_l2, l3_trunc, norms, two-layer gamma_t), training-mode dropout, the C=256 model,
# the energy/force head, the equivariant heads and the GEMM layout tests.  Run on the GPU box:
#   gpurun --timeout 3000 -- bash tools/gpu_sanitize.sh
# Logs land in gpurun_out/sanitizer_<tool>.txt (copy the summaries to profiles/).
set -u
mkdir -p gpurun_out
SEL='(golden_forward_backward and (cfg1 or yaml_l2 or l3_trunc or norms_l3 or eu_mlp- or eu_linw_ln_ev or eu_linwa_postln_gated or eu_mlpa_edgeln)) or attention_dropout or cfg2_width or (head_energy_forces_golden and l2) or dipole_and_spatial or (gemm_layouts and (257 or 1000)) or fp16_split'
TOOLS=${1:-"memcheck racecheck racecheck1cta synccheck initcheck"}
for tool in $TOOLS; do
  extra=""
  sel="$SEL"
  ncta=""
  # racecheck1cta: the same pass with the GEMM forced to single-CTA tiles (GOTEN_GEMM_NCTA=1): the only hazards the
  # default pass reports are on the tcgen05.alloc address slot of the CTA-PAIR kernels (written by the allocator of
  # the pair, ordered by the cluster barrier, which racecheck does not model) - this run shows every other kernel clean
  if [ "$tool" = racecheck1cta ]; then tool=racecheck; ncta=1; tag=racecheck_1cta; else tag=$tool; fi
  # initcheck instruments every global access (x100 slower): the two smallest golden cases only
  [ "$tool" = initcheck ] && sel='golden_forward_backward and (cfg1 or yaml_l2)'
  [ "$tool" = memcheck ] && extra="--leak-check no"
  start=$(date +%s)
  GOTEN_GEMM_NCTA=$ncta timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 40 --error-exitcode 9 \
    --log-file gpurun_out/sanitizer_${tag}_raw.txt \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$sel" -p no:cacheprovider > gpurun_out/sanitizer_${tag}_pytest.txt 2>&1
  rc=$?
  end=$(date +%s)
  {
    echo "# compute-sanitizer --tool $tool ${ncta:+(GOTEN_GEMM_NCTA=1)}  (rc=$rc, $((end-start)) s)  selection: $sel"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|Barrier error|========= (Error|Warning)" gpurun_out/sanitizer_${tag}_raw.txt | sed 's/+0x[0-9a-f]*//' | cut -c1-200 | sort | uniq -c | sort -rn | head -24
    echo "# pytest tail:"
    tail -4 gpurun_out/sanitizer_${tag}_pytest.txt
  } > gpurun_out/sanitizer_${tag}.txt
  head -c 20000 gpurun_out/sanitizer_${tag}_raw.txt > gpurun_out/sanitizer_${tag}_head.txt
  cat gpurun_out/sanitizer_${tag}.txt
done
