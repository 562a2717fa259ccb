"""Condense an ncu CSV launch list of the GEMM kernels of one training step into profiles/ncu_gemm_summary.json.

Input: the --csv log of
  ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,\
dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv3x3_kernel|igemm_|wgrad3x3_kernel|\
stem_fprop_kernel|stem_wgrad_kernel" python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline
Only the launches of the LAST step are used (steps are delimited by stem_fprop_kernel). bench.py reads the result
for `roofline.traffic` / `tensor_pipe_active_pct_ncu` / `conv_tc_util_pct_flop_weighted_ncu`."""
import collections
import csv
import json
import re
import sys


def to_unit(value, unit):
    v = float(value.replace(",", ""))
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}
    return v * scale.get(unit, 1.0)


def main(src, dst):
    rows = list(csv.reader(open(src, newline="")))
    hdr = next(r for r in rows if "Kernel Name" in r)
    ix = {k: i for i, k in enumerate(hdr)}
    launches = collections.OrderedDict()  # launch id -> {name, metrics}
    for r in rows:
        if len(r) != len(hdr) or r is hdr or r[ix["ID"]] == "ID":
            continue
        d = launches.setdefault(r[ix["ID"]], {"name": re.sub(r"^void\s+", "", re.sub(r"(<|\().*", "", r[ix["Kernel Name"]])).replace("qt::", ""),
                                              "full": r[ix["Kernel Name"]]})
        try:
            d[r[ix["Metric Name"]]] = to_unit(r[ix["Metric Value"]], r[ix["Metric Unit"]])
        except ValueError:
            pass
    seq = list(launches.values())
    starts = [i for i, l in enumerate(seq) if l["name"] == "stem_fprop_kernel"]
    last = seq[starts[-1]:] if starts else seq
    # the backward kernels of the last step follow its forward; everything after the last stem_fprop belongs to it
    fam = collections.OrderedDict()
    tot_t = tot_w = 0.0
    for l in last:
        t = l.get("gpu__time_duration.sum", 0.0)
        a = l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
        b = l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
        f = fam.setdefault(l["name"], {"launches_per_step": 0, "us_per_step": 0.0, "_bytes": 0.0, "_act": 0.0})
        f["launches_per_step"] += 1
        f["us_per_step"] += t
        f["_bytes"] += b
        f["_act"] += a * t
        tot_t += t
        tot_w += a * t
    out = collections.OrderedDict()
    for k, f in fam.items():
        out[k] = {"launches_per_step": f["launches_per_step"], "us_per_step": round(f["us_per_step"], 1),
                  "dram_bytes_per_launch": int(f["_bytes"] / f["launches_per_step"]),
                  "tensor_active_pct": round(f["_act"] / max(f["us_per_step"], 1e-9), 2)}
    out["flop_weighted_tensor_active_pct"] = round(tot_w / max(tot_t, 1e-9), 2)
    out["gemm_us_per_step"] = round(tot_t, 1)
    out["note"] = ("one training step (B=256) under ncu --clock-control none; sm__pipe_tensor_cycles_active.avg."
                   "pct_of_peak_sustained_elapsed per launch, weighted by launch duration (MMA work ~ tensor-active time, so "
                   "this is the FLOP-weighted pipe utilisation up to the border-pixel work of the slab kernels); "
                   "dram bytes = dram__bytes_read.sum + dram__bytes_write.sum")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "profiles/ncu_gemm_summary.json")
