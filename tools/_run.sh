mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^frame" | tail -40 | grep -E "passed|failed|Error|assert|^E " > gpurun_out/s4_tests9.log
cat gpurun_out/s4_tests9.log
python bench.py --no-cpu-baseline > gpurun_out/s4_bench9_molpcba.log 2>&1
python bench.py --no-cpu-baseline --config code2 > gpurun_out/s4_bench9_code2.log 2>&1
python - <<'PY'
import json
for f in ["gpurun_out/s4_bench9_molpcba.log","gpurun_out/s4_bench9_code2.log"]:
    for l in open(f):
        if l.startswith("{"):
            j=json.loads(l); print(j["config"]["workload"][:30], round(j["value"]), round(j["ms_per_step"],3), j["e2e"]["value"], j["with_optimizer"], j["gpu_launches"])
            for k in j["roofline_kernels"]: print("   ", k["kernel"][:40], round(k["achieved"],1), round(k["frac"],4), round(k["avg_launch_us"],1), k["share_of_step"])
PY
