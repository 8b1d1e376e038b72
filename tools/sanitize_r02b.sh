#!/bin/bash
# compute-sanitizer over the kernels added late in round 2: k_agg_bwd4 (mbarrier / TMEM / tcgen05 inside the adjoint),
# k_agg_bwd3p (packed mask), the warp-per-graph pooled-query kernels, gt_relu_bwd_colsum, the bf16 residual operand in the
# GEMM's TMA-store epilogue.      bash tools/sanitize_r02b.sh [outdir]
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SEL='fused_in_adjoint or packed_mask or pooled_query_attention or relu_bwd_colsum or tcgen05_forced'
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  log="$OUT/r02_sanitizer_${tool}_v2.txt"
  echo "== compute-sanitizer --tool $tool : pytest tests/test_ops_gpu.py tests/test_round2_gpu.py -k '$SEL'" > "$log"
  timeout 300 $CS --tool $tool --print-limit 30 --error-exitcode 9 \
      python -m pytest tests/test_ops_gpu.py tests/test_round2_gpu.py -x -q -m gpu -k "$SEL" -p no:cacheprovider >> "$log" 2>&1
  echo "exit code: $?" >> "$log"
  tail -4 "$log"
done
