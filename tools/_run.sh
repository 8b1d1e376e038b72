# scratch command file for `gpurun -- 'bash tools/_run.sh'` (what the last GPU call of the session ran): full GPU test suite,
# smoke, the headline bench; outputs under gpurun_out/ (copy what should be kept into profiles/)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^frame" | tail -40 | grep -E "passed|failed|Error|assert|^E " > gpurun_out/tests_gpu.log
cat gpurun_out/tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_molpcba.log 2>&1
