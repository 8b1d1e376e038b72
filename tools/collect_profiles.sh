#!/bin/bash
# Run on the GPU box (gpurun): ncu launch lists + full captures of the hot kernels for configs molpcba and code2.
# Outputs go to gpurun_out/; summarise them on the build box with tools/ncu_summary.py / tools/ncu_launch_summary.py.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e --no-optimizer"
for cfg in molpcba code2; do
  ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 4000 -c 700 --csv \
      --log-file gpurun_out/launches_${cfg}.csv $B --config $cfg > gpurun_out/ncu_launch_${cfg}.log 2>&1
  ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:k_agg_fwd -s 10 -c 2 -f \
      -o gpurun_out/prof_agg_fwd_${cfg} $B --config $cfg > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:k_agg_bwd -s 10 -c 2 -f \
      -o gpurun_out/prof_agg_bwd_${cfg} $B --config $cfg > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:k_mha_ -s 24 -c 5 -f \
      -o gpurun_out/prof_mha_${cfg} $B --config $cfg > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:k_gemm_tc -s 200 -c 6 -f \
      -o gpurun_out/prof_gemm_${cfg} $B --config $cfg > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
