mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^frame" | tail -40 | grep -E "passed|failed|Error|assert" > gpurun_out/s4_tests8.log
cat gpurun_out/s4_tests8.log
timeout 120 python tools/mha_bench.py code2 molpcba 2>&1 | grep "impl=2" > gpurun_out/s4_mha_bench8.log
cat gpurun_out/s4_mha_bench8.log
timeout 120 python tools/graph_trace.py code2 --raw > gpurun_out/s4_trace8_code2.log 2>&1
head -2 gpurun_out/s4_trace8_code2.log
