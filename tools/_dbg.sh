#!/bin/bash
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e --no-optimizer --no-extra-configs"
NCU="ncu --clock-control none --graph-profiling node"
timeout 300 $NCU --set full --import-source on -k regex:"k_agg_bwd3p|k_mha_cls_" -s 6 -c 4 -f -o gpurun_out/r02_prof_agg3p_syn $B --config syn > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02_prof_agg3p_syn.ncu-rep > gpurun_out/r02_ncu_agg3p_syn_v2.txt 2>&1
rm -f gpurun_out/*.ncu-rep
grep -E "^--|duration|warp instr|dram %|warps active|stalls" gpurun_out/r02_ncu_agg3p_syn_v2.txt | cut -c1-150 | head -24
