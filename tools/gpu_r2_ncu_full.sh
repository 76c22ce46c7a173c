#!/bin/bash
# ncu --set full of every library kernel of one hot-path pass + the library's scope names, converted to JSON on the box
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 300 python tools/hot_path_once.py --tier-b --sequence gpurun_out/seq_$TAG.json > gpurun_out/plain3_$TAG.log 2>&1 || { tail -20 gpurun_out/plain3_$TAG.log; exit 1; }
NPASS=$(python -c "import json; print(len(json.load(open('gpurun_out/seq_$TAG.json'))))")
echo "scopes per pass: $NPASS"
REP=/tmp/prof_full_$TAG
# skip the first pass (warm-up) of the library's kernels: -s NPASS launches matching the regex, capture the second pass
timeout 1800 ncu --set full --clock-control none -k regex:"umma_kernel|lazy_|aspp_|fcd_|sgd_step|adam_step|upsample|fast_hist|gan_loss|nchw_|nhwc_|conv1_fwd|conv2_dgrad|conv1_dgrad|per_class" -s $NPASS -c $NPASS -o $REP python tools/hot_path_once.py --tier-b > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"; tail -2 gpurun_out/ncu_full_$TAG.log
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/full_${TAG}_raw.csv 2> /dev/null
python tools/ncu_full_to_json.py gpurun_out/full_${TAG}_raw.csv gpurun_out/seq_$TAG.json gpurun_out/${TAG}_hot_kernels_ncu_full.json gpurun_out/${TAG}_ncu_traffic.json
