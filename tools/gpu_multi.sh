#!/bin/bash
# bench at N GPUs of one box (weak scaling): $1 = N
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --also-trunk-bf16 0 > gpurun_out/final_bench_${N}gpu.json 2> gpurun_out/final_bench_${N}gpu.err
echo "bench $N gpus exit $?"
tail -n 3 gpurun_out/final_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/final_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['e2e']['value'], d['config']['parallelism'])
PY
