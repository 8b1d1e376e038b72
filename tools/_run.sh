# scratch command file for `gpurun -- 'bash tools/_run.sh'` (what the last GPU call of the session ran)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "passed|failed|Error|^E " | tail -3
python bench.py > gpurun_out/r01_bench_molpcba_v9.log 2>&1
python bench.py --config syn --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-optimizer > gpurun_out/r01_bench_syn_v9b.log 2>&1
