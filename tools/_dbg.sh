#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/graph_trace.py syn > gpurun_out/r02_graph_trace_syn_n.txt 2>&1; grep -E "us/step|onehot|bn_bwd_apply|aggregate_bwd" gpurun_out/r02_graph_trace_syn_n.txt | cut -c1-140
