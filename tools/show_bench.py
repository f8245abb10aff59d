import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        l = l.strip()
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        print(d["config"]["workload"], d["dtype"], "%.3f ms" % d["ms_per_step"], "%.3e DoF/s" % d["value"], "frac %.3f" % d["roofline"]["frac"])
        for k, v in (d.get("sweep") or {}).items():
            print("  ", k, "%.3f ms %.3e DoF/s frac %.3f" % (v["ms"], v["dofs_per_s"], v["frac"]))
