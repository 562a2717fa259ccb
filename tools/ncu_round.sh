#!/bin/bash
# ncu evidence of one round (run under gpurun on one B200): launch list of the headline step, GEMM-kernel metrics for
# tools/ncu_summary.py, and --set full captures of the dominant conv kernel and the quadtree-stage kernels.
# usage: tools/ncu_round.sh r02
tag=${1:-r02}
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline"
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_ncu_launch_list_step.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"conv3x3_kernel|igemm_|wgrad3x3_kernel|stem_fprop_kernel|stem_wgrad_kernel" --csv \
    --log-file gpurun_out/${tag}_ncu_gemm_launches.csv $CMD > gpurun_out/${tag}_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"quadtree_pool|head_tail|adam_multi|bn_apply|bn_bwd|bn_relu_maxpool|stem_bn_pool|stem_pack" -s 60 -c 80 --csv \
    --log-file gpurun_out/${tag}_ncu_stream_launches.csv $CMD > gpurun_out/${tag}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s 30 -c 1 -o gpurun_out/${tag}_conv3x3_full $CMD > gpurun_out/${tag}_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:quadtree_pool -s 6 -c 2 -o gpurun_out/${tag}_quadtree_pool_full $CMD > gpurun_out/${tag}_ncu5.log 2>&1
ls -la gpurun_out/${tag}_*
