mkdir -p gpurun_out
python tools/agg_bench.py 2>&1 | grep -v "^  " > gpurun_out/s4_agg_v3c.log
GT_AGG_VARIANT=2 python tools/agg_bench.py 2>&1 | grep -v "^  " > gpurun_out/s4_agg_v2c.log
cat gpurun_out/s4_agg_v3c.log gpurun_out/s4_agg_v2c.log
