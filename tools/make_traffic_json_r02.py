"""profiles/ncu_traffic.json from the text summaries tools/ncu_summary.py wrote (profiles/r02_ncu_*_<tag>.txt): DRAM bytes
(read + write) per launch of the hot kernels, averaged over the captured launches.  bench.py copies them into
roofline.traffic.    python tools/make_traffic_json_r02.py v3"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "v3"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(path):
    """-> [(kernel name, dram bytes)]"""
    if not os.path.exists(path):
        return []
    res, name, rd = [], None, None
    for line in open(path):
        if line.startswith("-- "):
            name, rd = line[3:].strip(), None
        m = re.match(r"\s+dram (read|write)\s+([0-9.]+) (\w+)", line)
        if m and name:
            v = float(m.group(2)) * UNIT.get(m.group(3), 1.0)
            if m.group(1) == "read":
                rd = v
            elif rd is not None:
                res.append((name, rd + v))
    return res


def mean(rows, *subs):
    x = [b for n, b in rows if not subs or any(s in n for s in subs)]
    return sum(x) / len(x) if x else None


P = lambda stem, cfg: launches(os.path.join(ROOT, "profiles", f"r02_ncu_{stem}_{cfg}_{TAG}.txt"))
note = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, cold cache (profiles/r02_ncu_*_%s.txt)" % TAG
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
out = json.load(open(path)) if os.path.exists(path) else {}
out["syn"] = {"aggregate": mean(P("agg", "syn"), "k_agg_"), "mha": mean(P("mha", "syn"), "k_mha_tc", "k_mha_delta"),
              "mha_local": None, "gemm": mean(P("gemm", "syn"), "k_gemm_tc"), "note": note}
m = dict(out.get("molpcba", {}))
if mean(P("agg", "molpcba"), "k_agg_") is not None:
    m["aggregate"] = mean(P("agg", "molpcba"), "k_agg_")
out["molpcba"] = m
c = dict(out.get("code2", {}))
if mean(P("mha", "code2"), "k_mha_tc") is not None:
    c["mha"] = mean(P("mha", "code2"), "k_mha_tc", "k_mha_delta")
out["code2"] = c
out["code2-pna"] = {"aggregate": mean(P("pna", "code2-pna"), "k_pna_"), "mha": c.get("mha"), "mha_local": None,
                    "gemm": c.get("gemm"), "note": note + "; mha / gemm figures of the code2 capture (same transformer)"}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
