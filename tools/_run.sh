mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^frame" | tail -40 | grep -E "passed|failed|Error|assert|^E " > gpurun_out/r01_tests_v8.log
cat gpurun_out/r01_tests_v8.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01_smoke_v8.log 2>&1; tail -3 gpurun_out/r01_smoke_v8.log
python bench.py > gpurun_out/r01_bench_molpcba_v8.log 2>&1
python bench.py --config code2 --no-cpu-baseline > gpurun_out/r01_bench_code2_v8.log 2>&1
