#!/bin/bash
# one development iteration on the B200 box: the tensor-core suites, the step test, then the bench.
mkdir -p gpurun_out
rm -f gpurun_out/iter_summary.txt
for t in umma aspp fcd step; do
  timeout 600 python -m pytest tests/test_gpu_$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/iter_$t.log 2>&1
  echo "test_$t exit $?" >> gpurun_out/iter_summary.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err
echo "bench exit $?" >> gpurun_out/iter_summary.txt
cat gpurun_out/iter_summary.txt
tail -n 5 gpurun_out/iter_umma.log gpurun_out/iter_aspp.log gpurun_out/iter_step.log
cat gpurun_out/iter_bench.json
