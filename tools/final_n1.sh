#!/bin/bash
# Final single-GPU lines of a round: headline (default flags), the secondary workloads and the reference arm.
tag=${1:-r02}
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
for w in attention_infer quadtree3d_train cnn_lstm_train; do
  python bench.py --workload $w > gpurun_out/${tag}_${w}_n1.json 2> gpurun_out/${tag}_${w}_n1.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
for f in bench_n1 attention_infer_n1 quadtree3d_train_n1 cnn_lstm_train_n1 bench_reference; do
  python - "$f" "$tag" <<'PY'
import json, sys
f, tag = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/{tag}_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"], 1), d["unit"], round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "launches", d.get("gpu_launches"),
          "roofline", (d.get("roofline") or {}).get("frac"), "clocks", d.get("clocks"))
except Exception as e:
    print(f, "unreadable:", e)
PY
done
