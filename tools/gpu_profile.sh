#!/bin/bash
# ncu evidence (run under gpurun, 1 GPU).  $1 = tag for the output names.
#  1. launch list of bench.py (every kernel of 2 eager iterations with its device time)
#  2. --set full capture of the library's kernels in the short hot-path script
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-kernel-events --cuda-graph 0 --cudnn-benchmark 0"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list exit $?"
python tools/hot_path_once.py > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"umma_kernel|lazy_|aspp_|fcd_|ce_kernel|ce_finalize|softmax_kernel|sgd_step|adam_step|upsample_|fast_hist|gan_loss|nchw_|nhwc_" -s 75 -c 90 -o gpurun_out/prof_hot_$TAG python tools/hot_path_once.py > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | grep $TAG
