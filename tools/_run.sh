python tools/mha_bench.py molpcba nci1 > gpurun_out/mha_bench2.log 2>&1
python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/t34.log
python bench.py --no-cpu-baseline > gpurun_out/bench27.log 2>&1
python bench.py --no-cpu-baseline --config nci1 > gpurun_out/bench27_nci1.log 2>&1
tail -5 gpurun_out/t34.log
