"""Print the headline and per-family roofline numbers of bench.py JSON lines side by side."""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "img/s", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]) if d.get("e2e") and d["e2e"]["value"] else None,
          "launches", d.get("gpu_launches"))
    r = d.get("roofline") or {}
    for k, v in sorted((r.get("per_family") or {}).items()):
        print("    %-44s %6.0f TFLOP/s %7.3f ms  x%g" % (k, v["tflops"], v["ms_per_step"], v["launches_per_step"]))
    for k, v in sorted((r.get("hbm_families") or {}).items()):
        print("    %-44s %6.0f GB/s    %7.3f ms  x%g  (%.2f of peak)" % (k, v["GBps"], v["ms_per_step"], v["calls_per_step"], v["frac_of_hbm_peak"]))
