mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^frame" | tail -40 > gpurun_out/s4_tests7.log
tail -25 gpurun_out/s4_tests7.log
timeout 120 python tools/graph_trace.py code2 --raw > gpurun_out/s4_trace7_code2.log 2>&1
head -2 gpurun_out/s4_trace7_code2.log
