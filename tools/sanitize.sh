#!/bin/bash
# compute-sanitizer passes over the hand-rolled mbarrier / TMEM / TMA kernels (SURVEY §4 sanitizer tier).
# Usage (on the GPU box): bash tools/sanitize.sh [outdir]     -> <outdir>/r02_sanitizer_{memcheck,racecheck}.txt
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SEL='tcgen05_forced or tcgen05_forward or tile_local_matches or pooled_query_attention or splitk_weight'
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  log="$OUT/r02_sanitizer_${tool}.txt"
  echo "== compute-sanitizer --tool $tool : pytest tests/test_ops_gpu.py tests/test_round2_gpu.py -k '$SEL'" > "$log"
  timeout 480 $CS --tool $tool --print-limit 30 --error-exitcode 9 \
      python -m pytest tests/test_ops_gpu.py tests/test_round2_gpu.py -x -q -m gpu -k "$SEL" -p no:cacheprovider >> "$log" 2>&1
  echo "exit code: $?" >> "$log"
  tail -5 "$log"
done
