#!/bin/bash
# end-of-round check on one B200: the whole GPU suite, smoke, the bench (both arms)
mkdir -p gpurun_out
rm -f gpurun_out/final_summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/final_gpu.txt 2>&1
timeout 1200 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/final_tests.log 2>&1
echo "pytest -m gpu exit $?" >> gpurun_out/final_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/final_summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err
echo "bench exit $?" >> gpurun_out/final_summary.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
echo "bench --impl reference exit $?" >> gpurun_out/final_summary.txt
cat gpurun_out/final_summary.txt
tail -n 4 gpurun_out/final_tests.log
tail -n 3 gpurun_out/final_smoke.log
cat gpurun_out/final_bench_ref.json | cut -c1-600
python tools/bench_diff.py gpurun_out/final_bench_1gpu.json > gpurun_out/final_diff.txt 2>&1; head -5 gpurun_out/final_diff.txt
