#!/bin/bash
# ncu launch list of the bench command (eager, sequential schedule so that every kernel is listed once per call)
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference --no-e2e --no-kernel-events --also-trunk-bf16 0 --cuda-graph 0 --overlap 0 --cudnn-benchmark 0"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -20 gpurun_out/plain_$TAG.log; exit 1; }
timeout 2400 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list exit $?"
python tools/summarize_ncu.py gpurun_out/launches_$TAG.csv gpurun_out/launches_$TAG.md | head -40
