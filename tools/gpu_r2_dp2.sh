#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/dp_bucket_check.py > gpurun_out/dp_check.log 2>&1
echo "check exit $?"; grep -E "DP_BUCKET|Error" gpurun_out/dp_check.log | head; cp gpurun_out/dp_check.log gpurun_out/dp_check_final.log
for cfg in "ASN_DIAG_SKIP_ALLREDUCE=1 ASN_DIAG_BUCKETS=0" "ASN_X=1" "ASN_BUCKETED_ALLREDUCE=0"; do
env $cfg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --also-trunk-bf16 0 --no-kernel-events --no-e2e > gpurun_out/dp2.json 2> gpurun_out/dp2.err
echo "[$cfg] exit $?"; tail -n 1 gpurun_out/dp2.err | cut -c1-200; python -c "
import json; d=json.loads(open('gpurun_out/dp2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'])"
done
