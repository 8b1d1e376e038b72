mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^frame" | tail -40 | grep -E "passed|failed|Error|assert|^E " > gpurun_out/s4_tests14.log
cat gpurun_out/s4_tests14.log
python tools/graph_trace.py molpcba 2>&1 | head -1
