mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^frame" | tail -40 | grep -E "passed|failed|Error|assert|^E " > gpurun_out/tests_gpu.log
cat gpurun_out/tests_gpu.log
for v in 1 0; do echo "tabgemm=$v $(GT_TABLE_GRAD_GEMM=$v python tools/graph_trace.py molpcba 2>&1 | head -1 | cut -c1-60)"; done
for v in 1 0; do echo "syn tabgemm=$v $(GT_TABLE_GRAD_GEMM=$v python bench.py --config syn --steps 6 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e --no-optimizer 2>&1 | grep -o '"ms_per_step": [0-9.]*')"; done
