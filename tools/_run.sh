mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^frame" | tail -40 | grep -E "passed|failed|Error|assert|^E " > gpurun_out/s4_tests13.log
cat gpurun_out/s4_tests13.log
for i in 1 2; do
python tools/graph_trace.py molpcba > gpurun_out/s4_trace13_molpcba.log 2>&1; grep -E "plain graph|aggregate_[fb]" gpurun_out/s4_trace13_molpcba.log
done
