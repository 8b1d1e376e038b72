#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in molpcba; do timeout 200 python tools/graph_trace.py $c > gpurun_out/r02_graph_trace_${c}_q.txt 2>&1; grep -E "us/step|gt_aggregate" gpurun_out/r02_graph_trace_${c}_q.txt | cut -c1-140; done
