#!/bin/bash
# Scaling lines on one 8-GPU box (run under `gpurun --gpus 8`): headline at 2/4/8 ranks, the 3-D and sequence models at 2/8.
tag=${1:-r02}   # usage: tools/scale_runs.sh <tag> [quick]
run() {  # workload nranks
  local wl=$1 n=$2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --workload $wl --steps 20 --warmup 5 --no-roofline > gpurun_out/${tag}_${wl}_n$n.json 2> gpurun_out/${tag}_${wl}_n$n.err
  tail -1 gpurun_out/${tag}_${wl}_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl', 'N=$n', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'e2e_fp32', round(d['e2e_fp32_inputs']['value']))"
}
if [ "$2" = "quick" ]; then  # one 8-rank line per workload (the full sweep costs ~40 GPU-minutes of an 8-GPU box)
  run quadtree_train 8; run quadtree3d_train 8; run cnn_lstm_train 8
  exit 0
fi
for n in 8 4 2; do run quadtree_train $n; done
for n in 8 2; do run quadtree3d_train $n; done
run cnn_lstm_train 8
