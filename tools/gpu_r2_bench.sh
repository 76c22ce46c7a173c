#!/bin/bash
# round-2: eval-path tests, the eval bench (config 5), the train bench with the GPU-eager comparator, the reference arm
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eval.py -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/test_eval_$TAG.log 2>&1
echo "pytest eval exit $?"; tail -n 5 gpurun_out/test_eval_$TAG.log
timeout 900 python bench.py --mode eval --frames 500 > gpurun_out/bench_eval_$TAG.json 2> gpurun_out/bench_eval_$TAG.err
echo "bench eval exit $?"; tail -c 1500 gpurun_out/bench_eval_$TAG.err; head -c 3000 gpurun_out/bench_eval_$TAG.json; echo
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    p=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print(p["ms_per_step"], p["hot_path_ms_per_step"], p["roofline"]["kernel"], p["roofline"]["frac"])
    print(json.dumps(p.get("reference_gpu_eager"), indent=1))
    print(p.get("cpu_baseline"))
except Exception as e: print("no bench line", e)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
echo "bench ref exit $?"; tail -c 800 gpurun_out/bench_ref_$TAG.err; head -c 1500 gpurun_out/bench_ref_$TAG.json
