#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/iter_summary.txt
timeout 400 python -m pytest tests/test_gpu_step.py tests/test_gpu_aspp.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/iter_step.log 2>&1
echo "test_step+aspp exit $?" >> gpurun_out/iter_summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err
echo "bench exit $?" >> gpurun_out/iter_summary.txt
cat gpurun_out/iter_summary.txt
grep -E "^E  |passed|failed|FAILED" gpurun_out/iter_step.log | cut -c1-300 | head -20
tail -n 3 gpurun_out/iter_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/iter_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','hot_path_ms_per_step')}, d['e2e']['value'], d.get('trunk_bf16_autocast'))
print(d['roofline'])
PY
