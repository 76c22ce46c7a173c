#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/iter_summary.txt
timeout 600 python -m pytest tests/test_gpu_lazy.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/iter_lazy.log 2>&1
echo "test_lazy exit $?" >> gpurun_out/iter_summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tier B > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err
echo "bench B exit $?" >> gpurun_out/iter_summary.txt
cat gpurun_out/iter_summary.txt
grep -E "^E  |passed|failed|FAILED" gpurun_out/iter_lazy.log | cut -c1-300 | head -30
tail -n 3 gpurun_out/iter_bench.err
python tools/bench_diff.py gpurun_out/iter_bench.json | head -45
