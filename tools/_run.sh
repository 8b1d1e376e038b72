mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^frame" | tail -30 > gpurun_out/s4_tests6.log
tail -3 gpurun_out/s4_tests6.log
timeout 120 python tools/graph_trace.py molpcba --raw > gpurun_out/s4_trace6_molpcba.log 2>&1
timeout 120 python tools/graph_trace.py code2 --raw > gpurun_out/s4_trace6_code2.log 2>&1
head -2 gpurun_out/s4_trace6_molpcba.log; head -2 gpurun_out/s4_trace6_code2.log
python bench.py --no-cpu-baseline > gpurun_out/s4_bench6_molpcba.log 2>&1
tail -c 300 gpurun_out/s4_bench6_molpcba.log
