#!/bin/bash
# A/B of environment switches on the per-kernel table: usage gpu_r2_ab.sh TAG "ENV1=.. ENV2=.." ["ENV..." ...]
TAG=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --also-trunk-bf16 0 --no-e2e > gpurun_out/ab_${TAG}_$i.json 2> gpurun_out/ab_${TAG}_$i.err
  echo "== [$envs] exit $?"; tail -c 300 gpurun_out/ab_${TAG}_$i.err
  python - <<PY
import json
try:
    p=json.loads(open("gpurun_out/ab_${TAG}_$i.json").read().strip().splitlines()[-1])
    print("ms/step", round(p["ms_per_step"],3), "hot", round(p["hot_path_ms_per_step"],3))
    print(" ".join(f"{k}={v['ms_per_step']:.3f}" for k,v in p["hot_path_kernels"].items() if k.startswith(("fcd_conv","aspp_"))))
except Exception as e: print("no line", e)
PY
done
