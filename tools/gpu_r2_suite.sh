#!/bin/bash
# round-2 GPU check: the whole GPU suite (optionally with extra env, e.g. ASN_HALO2=1), smoke, one bench line.
# usage: tools/gpu_r2_suite.sh TAG [ENV=VAL ...]
TAG=${1:-r02}; shift
for kv in "$@"; do export "$kv"; done
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/gpu_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/test_$TAG.log 2>&1
echo "pytest exit $?" | tee gpurun_out/summary_$TAG.txt
tail -n 15 gpurun_out/test_$TAG.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary_$TAG.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?" | tee -a gpurun_out/summary_$TAG.txt
tail -c 600 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    p=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print(p["ms_per_step"], p["hot_path_ms_per_step"], p["roofline"]["kernel"], p["roofline"]["frac"])
    for k,v in p["hot_path_kernels"].items(): print(f"{k:30s} {v['ms_per_step']:.4f} {v['frac']}")
except Exception as e: print("no bench line", e)
PY
