#!/bin/bash
# full GPU suite + ncu launch list of the trunk-less hot path + a short bench line.  usage: tools/gpu_r3_check.sh TAG
TAG=${1:-c1}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/test_$TAG.log 2>&1
echo "pytest exit $?"; tail -n 4 gpurun_out/test_$TAG.log; grep -E "^E  |FAILED" gpurun_out/test_$TAG.log | cut -c1-300 | head
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/glue_launches_$TAG.csv \
  python tools/glue_bench.py --reps 2 > gpurun_out/glue_ncu_$TAG.log 2>&1
echo "ncu exit $?"
python tools/summarize_ncu.py gpurun_out/glue_launches_$TAG.csv gpurun_out/glue_launches_$TAG.md > /dev/null; head -70 gpurun_out/glue_launches_$TAG.md | cut -c1-160
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --also-trunk-bf16 0 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -c 300 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    p=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print(p["ms_per_step"], "e2e", p["e2e"]["value"], "hot", p["hot_path_ms_per_step"], p["roofline"]["kernel"], p["roofline"]["frac"], "launches", p["gpu_launches"])
    for k,v in p["hot_path_kernels"].items(): print(f"{k:30s} {v['ms_per_step']:.4f} x{v['launches_per_step']:.0f} {v['frac']}")
except Exception as e: print("no bench line", e)
PY
