#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "layernorm or golden or parity or ln" 2>&1 | tail -2
for c in syn; do
timeout 200 python tools/graph_trace.py $c > gpurun_out/r02_graph_trace_${c}_p.txt 2>&1; grep -E "us/step|gt_layernorm_fwd \[530|gt_aggregate" gpurun_out/r02_graph_trace_${c}_p.txt | cut -c1-140
done
