#!/bin/bash
# Run on the GPU box (gpurun): ncu launch list + full captures of the hot kernels (round 2).
#   bash tools/collect_profiles_r02.sh [tag]     -> gpurun_out/r02_*<tag>*; summarise here with tools/ncu_summary.py
set -u
TAG=${1:-a}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e --no-optimizer --no-extra-configs"
NCU="ncu --clock-control none --graph-profiling node"
# launch list of the headline workload (cold-cache, serialised: compare SHARES, not absolutes)
$NCU --metrics gpu__time_duration.sum -s 700 -c 420 --csv --log-file gpurun_out/r02_launches_syn_${TAG}.csv $B --config syn > gpurun_out/r02_ncu_launch_syn_${TAG}.log 2>&1
$NCU --metrics gpu__time_duration.sum -s 1100 -c 640 --csv --log-file gpurun_out/r02_launches_molpcba_${TAG}.csv $B --config molpcba > gpurun_out/r02_ncu_launch_molpcba_${TAG}.log 2>&1
$NCU --metrics gpu__time_duration.sum -s 1000 -c 600 --csv --log-file gpurun_out/r02_launches_code2_${TAG}.csv $B --config code2 > gpurun_out/r02_ncu_launch_code2_${TAG}.log 2>&1
cap() {  # cfg name regex skip count
  $NCU --set full --import-source on -k regex:$3 -s $4 -c $5 -f -o gpurun_out/r02_prof_$2_$1_${TAG} $B --config $1 > /dev/null 2>&1
}
cap syn agg "k_agg_(fwd|bwd)3" 12 2
cap syn rowops "k_layernorm_bwd|k_layernorm_fwd|k_bn_bwd_apply|k_bn_bwd_reduce|k_bn_norm_fwd|k_colsum|k_relu_bwd" 60 8
cap syn mha "k_mha_" 20 6
cap syn gemm "k_gemm_tc" 60 8
cap code2-pna pna "k_pna_" 12 2
cap molpcba agg "k_agg_(fwd|bwd)3" 15 2
cap code2 mha "k_mha_tc" 24 5
# the .ncu-rep files stay on the box (gpurun copies back <= 64 MiB): summarise them here
for f in gpurun_out/r02_prof_*_${TAG}.ncu-rep; do
  python tools/ncu_summary.py $f > ${f%.ncu-rep}.txt 2>&1
done
python tools/ncu_launch_summary.py gpurun_out/r02_launches_syn_${TAG}.csv > gpurun_out/r02_launches_syn_${TAG}.txt 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02_launches_molpcba_${TAG}.csv > gpurun_out/r02_launches_molpcba_${TAG}.txt 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02_launches_code2_${TAG}.csv > gpurun_out/r02_launches_code2_${TAG}.txt 2>&1
rm -f gpurun_out/*.ncu-rep gpurun_out/r02_launches_*_${TAG}.csv
ls -la gpurun_out | tail -30
