for m in 0 2 10 15 8 3; do
GT_BN_SLAB=$m python bench.py --no-cpu-baseline --no-roofline --no-e2e > gpurun_out/bench25_slab$m.log 2>&1
done
python tools/gemm_bench.py > gpurun_out/gemm_bench6.log 2>&1
