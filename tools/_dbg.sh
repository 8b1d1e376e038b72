#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q -k "mha or golden" 2>&1 | tail -2
for pf in 0 296 0 296; do
GT_MHA_PREFETCH=$pf timeout 200 python tools/graph_trace.py syn > gpurun_out/r02_graph_trace_syn_k$pf.txt 2>&1; echo "pf=$pf"; grep -E "us/step|gt_mha_bwd|gt_mha_fwd" gpurun_out/r02_graph_trace_syn_k$pf.txt | cut -c1-120
done
