#!/bin/bash
TAG=${1:-ov}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_optim.py tests/test_gpu_dropin_scripts.py -q -m gpu --no-header -p no:cacheprovider -x 2>&1 | tail -15
for ov in 0 1; do
  timeout 600 python bench.py --steps 20 --warmup 5 --overlap $ov --no-cpu-baseline --no-gpu-reference --also-trunk-bf16 0 --no-kernel-events > gpurun_out/bench_${TAG}_$ov.json 2> gpurun_out/bench_${TAG}_$ov.err
  echo "overlap=$ov exit $?"; tail -c 400 gpurun_out/bench_${TAG}_$ov.err; cat gpurun_out/bench_${TAG}_$ov.json | cut -c1-300
done
