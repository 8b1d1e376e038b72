mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^frame" | tail -40 > gpurun_out/s4_tests3.log
python tools/graph_trace.py molpcba --raw > gpurun_out/s4_trace3_molpcba.log 2>&1
python tools/graph_trace.py code2 --raw > gpurun_out/s4_trace3_code2.log 2>&1
tail -2 gpurun_out/s4_tests3.log; head -2 gpurun_out/s4_trace3_molpcba.log; head -2 gpurun_out/s4_trace3_code2.log
