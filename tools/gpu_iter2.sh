#!/bin/bash
# Tier-B bring-up: lazy tests, then both tiers through the bench
mkdir -p gpurun_out
rm -f gpurun_out/iter_summary.txt
timeout 600 python -m pytest tests/test_gpu_lazy.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/iter_lazy.log 2>&1
echo "test_lazy exit $?" >> gpurun_out/iter_summary.txt
for t in fcd step; do
  timeout 300 python -m pytest tests/test_gpu_$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/iter_$t.log 2>&1
  echo "test_$t exit $?" >> gpurun_out/iter_summary.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tier B > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err
echo "bench B exit $?" >> gpurun_out/iter_summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tier A > gpurun_out/iter_bench_A.json 2> gpurun_out/iter_bench_A.err
echo "bench A exit $?" >> gpurun_out/iter_summary.txt
cat gpurun_out/iter_summary.txt
tail -n 40 gpurun_out/iter_lazy.log
tail -n 3 gpurun_out/iter_fcd.log gpurun_out/iter_step.log
tail -n 5 gpurun_out/iter_bench.err
python tools/bench_diff.py gpurun_out/iter_bench.json gpurun_out/iter_bench_A.json
