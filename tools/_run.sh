mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^frame" | tail -40 | grep -E "passed|failed|Error|assert|^E " > gpurun_out/s4_tests12.log
cat gpurun_out/s4_tests12.log
timeout 120 python tools/graph_trace.py molpcba > gpurun_out/s4_trace12_molpcba.log 2>&1
timeout 120 python tools/graph_trace.py code2 > gpurun_out/s4_trace12_code2.log 2>&1
grep -E "plain graph" gpurun_out/s4_trace12_molpcba.log gpurun_out/s4_trace12_code2.log
