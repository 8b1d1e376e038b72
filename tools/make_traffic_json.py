"""profiles/ncu_traffic.json from the ncu captures of tools/collect_profiles.sh: DRAM bytes (read + write) per launch of
the hot kernels, averaged over the captured launches.  bench.py copies them into roofline.traffic."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dram_bytes(rep, name_filter=None):
    if not os.path.exists(rep):
        return []
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    H, U = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d, u = dict(zip(H, r)), dict(zip(H, U))
        if name_filter and name_filter not in d.get("Kernel Name", ""):
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(d[k])
            unit = u[k].lower()
            tot += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        res.append(tot)
    return res


def mean(x):
    return sum(x) / len(x) if x else None


out = {}
for cfg in ("molpcba", "code2"):
    g = lambda name, f=None: dram_bytes(os.path.join(ROOT, "gpurun_out", f"prof_{name}_{cfg}.ncu-rep"), f)
    agg = g("agg_fwd") + g("agg_bwd")
    out[cfg] = {"aggregate": mean(agg), "mha": mean(g("mha", "k_mha_tc") + g("mha", "k_mha_delta")),
                "mha_local": mean(g("mha", "k_mha_loc")), "gemm": mean(g("gemm")),
                "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, cold cache"}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
