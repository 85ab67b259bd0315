"""Pretty-print one bench.py JSON line: python tools/show_bench.py <file>"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 1), d["unit"], "| ms/step", round(d["ms_per_step"], 4), "| n_gpus", d["n_gpus"], "| launches", d.get("gpu_launches"))
print("e2e", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items()})
print("clocks", d.get("clocks"))
if d.get("roofline"):
    r = d["roofline"]; print("roofline", r["kernel"], round(r["achieved"], 1), "/", r["peak"], "=", round(r["frac"], 3), "traffic", r.get("traffic"))
print("cpu_baseline", d.get("cpu_baseline"))
ex = d.get("extras", {})
for k, v in ex.items():
    if k == "sharded" and isinstance(v, dict):
        for kk, vv in v.items():
            print("  sharded:", kk, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in vv.items() if a not in ("note",)} if isinstance(vv, dict) else vv)
    elif k == "comparators" and isinstance(v, dict):
        for kk, vv in v.items():
            print("  cmp:", kk, vv)
    else:
        print(" ", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()} if isinstance(v, dict) else v)
