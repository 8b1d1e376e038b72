mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01_smoke.log 2>&1; tail -4 gpurun_out/r01_smoke.log
for cfg in molpcba code2 syn code2-pna nci1; do
  python bench.py --config $cfg $( [ $cfg = molpcba ] || echo --no-cpu-baseline ) $( [ $cfg = syn ] && echo "--steps 8 --warmup 3" ) > gpurun_out/r01_bench_${cfg}_v6.log 2>&1
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference_molpcba_v6.log 2>&1
bash tools/collect_profiles.sh > gpurun_out/s4_collect.log 2>&1
for cfg in molpcba code2; do
  python tools/ncu_summary.py gpurun_out/prof_agg_fwd_${cfg}.ncu-rep gpurun_out/prof_agg_bwd_${cfg}.ncu-rep gpurun_out/prof_mha_${cfg}.ncu-rep gpurun_out/prof_gemm_${cfg}.ncu-rep > gpurun_out/r01_ncu_hot_kernels_${cfg}_v6.txt 2>&1
done
python tools/make_traffic_json.py > gpurun_out/ncu_traffic_v6.json 2>&1
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic.json
rm -f gpurun_out/*.ncu-rep
python tools/graph_trace.py molpcba > gpurun_out/r01_graph_trace_molpcba_v6.txt 2>&1
python tools/graph_trace.py code2 > gpurun_out/r01_graph_trace_code2_v6.txt 2>&1
du -sh gpurun_out
