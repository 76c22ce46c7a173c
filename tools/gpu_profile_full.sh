#!/bin/bash
# ncu --set full capture of the library's kernels in the short hot-path script (run under gpurun, 1 GPU).
# The .ncu-rep is summarised to CSV on the box (gpurun_out/ carries at most 64 MiB back) and kept only if small.
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 300 python tools/hot_path_once.py --tier-b > gpurun_out/plain2_$TAG.log 2>&1 || { tail -20 gpurun_out/plain2_$TAG.log; exit 1; }
REP=/tmp/prof_hot_$TAG
timeout 1500 ncu --set full --clock-control none -k regex:"umma_kernel|lazy_|aspp_|fcd_|sgd_step|adam_step|upsample_argmax|fast_hist|gan_loss|nchw_|nhwc_" -s 56 -c 60 -o $REP python tools/hot_path_once.py --tier-b > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/prof_hot_${TAG}_raw.csv 2> gpurun_out/ncu_export_$TAG.log
ls -la $REP.ncu-rep gpurun_out/prof_hot_${TAG}_raw.csv
SZ=$(stat -c %s $REP.ncu-rep)
if [ "$SZ" -lt 40000000 ]; then cp $REP.ncu-rep gpurun_out/; fi
tail -3 gpurun_out/ncu_full_$TAG.log
