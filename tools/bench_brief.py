#!/usr/bin/env python
"""Short view of a bench.py JSON line: python tools/bench_brief.py FILE"""
import json, sys
l = [x for x in open(sys.argv[1]) if x.startswith("{")][-1]
d = json.loads(l)
def row(name, m):
    k = m.get("kernels", {})
    s = " ".join("%s=%.0f" % (n.replace("mssvt_", "")[:14], 1000 * v["ms_per_step"]) for n, v in k.items())
    e = m.get("e2e", {})
    print("%-7s %.4f ms  e2e %.4f  | %s" % (name, m["ms_per_step"], e.get("ms_per_step", 0), s))
row(d["dtype"], d)
for n, m in d.get("modes", {}).items():
    row(n, m)
if "e2e_points" in d: print("e2e_points %.4f ms" % d["e2e_points"]["ms_per_step"])
print("parity", d.get("parity"))
