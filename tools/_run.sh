python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/t29.log
python bench.py --no-cpu-baseline > gpurun_out/bench21.log 2>&1
python bench.py --no-cpu-baseline --no-branch-stream --no-roofline --no-e2e > gpurun_out/bench21_nobranch.log 2>&1
python bench.py --no-cpu-baseline --config code2 > gpurun_out/bench21_code2.log 2>&1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 4000 -c 700 --csv --log-file gpurun_out/launches_molpcba_v6.csv $B --config molpcba > gpurun_out/ncu_launch_molpcba.log 2>&1
tail -5 gpurun_out/t29.log
