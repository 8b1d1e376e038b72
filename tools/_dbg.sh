#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for c in syn molpcba; do
timeout 200 python tools/graph_trace.py $c > gpurun_out/r02_graph_trace_${c}_l.txt 2>&1; grep -E "us/step|bn_bwd_reduce" gpurun_out/r02_graph_trace_${c}_l.txt | cut -c1-140
done
