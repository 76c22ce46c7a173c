#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/iter_summary.txt
for t in fcd step; do
  timeout 400 python -m pytest tests/test_gpu_$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/iter_$t.log 2>&1
  rc=$?
  echo "test_$t exit $rc" >> gpurun_out/iter_summary.txt
  if [ $rc -ne 0 ]; then grep -E "^E  |passed|failed|FAILED" gpurun_out/iter_$t.log | cut -c1-300 | head -20; cat gpurun_out/iter_summary.txt; exit 1; fi
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err
echo "bench exit $?" >> gpurun_out/iter_summary.txt
$AB_ENV timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/iter_bench_b.json 2> gpurun_out/iter_bench_b.err
echo "bench(B: $AB_ENV) exit $?" >> gpurun_out/iter_summary.txt
cat gpurun_out/iter_summary.txt
tail -n 3 gpurun_out/iter_bench.err
python tools/bench_diff.py gpurun_out/iter_bench.json gpurun_out/iter_bench_b.json > gpurun_out/iter_diff.txt 2>&1; head -48 gpurun_out/iter_diff.txt
