#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/iter_summary.txt
for t in optim step; do
  timeout 400 python -m pytest tests/test_gpu_$t.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/iter_$t.log 2>&1
  echo "test_$t exit $?" >> gpurun_out/iter_summary.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err
echo "bench exit $?" >> gpurun_out/iter_summary.txt
timeout 300 python tools/opt_probe.py > gpurun_out/opt_probe.json 2>/dev/null
cat gpurun_out/iter_summary.txt
grep -E "^E  |passed|failed|FAILED" gpurun_out/iter_optim.log gpurun_out/iter_step.log | cut -c1-300 | head -30
tail -n 3 gpurun_out/iter_bench.err
cat gpurun_out/opt_probe.json
python tools/bench_diff.py gpurun_out/iter_bench.json | head -12
