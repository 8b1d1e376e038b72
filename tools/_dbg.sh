#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/r02_tests_gpu_v3.txt; cat gpurun_out/r02_tests_gpu_v3.txt
F="--steps 30 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e --no-optimizer --no-extra-configs"
for c in molpcba code2; do timeout 100 python bench.py $F --config $c 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:20], d['value'], d['ms_per_step'], d['gpu_launches'])"; done
timeout 100 python tools/aten_probe.py molpcba 2>/dev/null | awk '{n+=$1} END {print "aten rows total count", n}'
