#!/bin/bash
# ncu evidence for bench.py (run under gpurun, 1 GPU).  $1 = tag for output names.
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:umma_kernel -s 40 -c 6 -o gpurun_out/prof_umma_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | tail -n 12
