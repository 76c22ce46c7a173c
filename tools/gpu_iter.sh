#!/bin/bash
# one development iteration on the B200 box: the tensor-core suites, the step test, then the bench
# (and, with AB=1, the bench again with CTA pairs switched off).
mkdir -p gpurun_out
rm -f gpurun_out/iter_summary.txt
for t in umma aspp fcd step; do
  timeout 300 python -m pytest tests/test_gpu_$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/iter_$t.log 2>&1
  rc=$?
  echo "test_$t exit $rc" >> gpurun_out/iter_summary.txt
  if [ $rc -ne 0 ]; then tail -n 30 gpurun_out/iter_$t.log; cat gpurun_out/iter_summary.txt; exit 1; fi
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err
echo "bench exit $?" >> gpurun_out/iter_summary.txt
if [ "$AB" = "1" ]; then
  ASN_PAIR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/iter_bench_nopair.json 2> gpurun_out/iter_bench_nopair.err
  echo "bench(no pair) exit $?" >> gpurun_out/iter_summary.txt
fi
cat gpurun_out/iter_summary.txt
tail -n 3 gpurun_out/iter_umma.log gpurun_out/iter_aspp.log gpurun_out/iter_fcd.log gpurun_out/iter_step.log
python tools/bench_diff.py gpurun_out/iter_bench.json gpurun_out/iter_bench_nopair.json 2>/dev/null
