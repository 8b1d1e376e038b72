mkdir -p gpurun_out
python bench.py > gpurun_out/r01_bench_molpcba_v8.log 2>&1
python bench.py --config code2-pna --no-cpu-baseline --no-optimizer > gpurun_out/r01_bench_code2-pna_v8.log 2>&1
python bench.py --config code2 --no-cpu-baseline > gpurun_out/r01_bench_code2_v8.log 2>&1
