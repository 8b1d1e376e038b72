#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/graph_trace.py syn > gpurun_out/r02_graph_trace_syn_m.txt 2>&1; grep -E "us/step|mha_cls" gpurun_out/r02_graph_trace_syn_m.txt | cut -c1-140
