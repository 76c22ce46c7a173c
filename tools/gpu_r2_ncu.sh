#!/bin/bash
# targeted ncu --set full capture: usage gpu_r2_ncu.sh TAG 'kernel-regex' [count]
TAG=$1; RX=$2; CNT=${3:-8}
mkdir -p gpurun_out
timeout 300 python tools/hot_path_once.py --tier-b > gpurun_out/plain_$TAG.log 2>&1 || { tail -20 gpurun_out/plain_$TAG.log; exit 1; }
REP=/tmp/prof_$TAG
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RX" -c $CNT -o $REP python tools/hot_path_once.py --tier-b > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_$TAG.log
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_raw.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_source.csv 2>/dev/null
ls -la $REP.ncu-rep gpurun_out/ncu_${TAG}_raw.csv gpurun_out/ncu_${TAG}_source.csv
SZ=$(stat -c %s $REP.ncu-rep); if [ "$SZ" -lt 30000000 ]; then cp $REP.ncu-rep gpurun_out/prof_$TAG.ncu-rep; fi
