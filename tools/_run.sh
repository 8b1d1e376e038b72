mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_parity_gpu.py -q -x 2>&1 | grep -E "passed|failed|Error|^E " | tail -5
python bench.py --config molpcba > gpurun_out/r01_bench_molpcba_v7.log 2>&1
python bench.py --config syn --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/r01_bench_syn_v7.log 2>&1
python bench.py --config code2 --no-cpu-baseline > gpurun_out/r01_bench_code2_v7.log 2>&1
tail -c 200 gpurun_out/r01_bench_molpcba_v7.log
