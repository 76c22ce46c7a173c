#!/bin/bash
# run-to-run variance of the headline number, cudnn.benchmark on / off
mkdir -p gpurun_out
for i in 1 2 3; do for cb in 1 0; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-kernel-events --also-trunk-bf16 0 --cudnn-benchmark $cb 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cudnn_benchmark=$cb run $i ms/step', round(d['ms_per_step'],3))"
done; done
