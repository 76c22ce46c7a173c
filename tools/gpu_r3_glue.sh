#!/bin/bash
# check of the rewritten layout kernels: the parity tests that cover them, then the trunk-less per-kernel bench
# (tools/glue_bench.py) once per environment in $CFGS (default: the new forms, then the round-1 forms with ASN_GLUE=0).
# usage: [CFGS="A=1 ASN_GLUE=0"] tools/gpu_r3_glue.sh TAG [test files...]
TAG=${1:-g1}; shift
TESTS=${@:-tests/test_gpu_aspp.py tests/test_gpu_fcd.py tests/test_gpu_lazy.py tests/test_gpu_step.py}
mkdir -p gpurun_out
timeout 600 python -m pytest $TESTS -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/glue_test_$TAG.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|FAILED|Error" gpurun_out/glue_test_$TAG.log | cut -c1-300 | head -30
for cfg in ${CFGS:-A=1 ASN_GLUE=0}; do
  env $cfg timeout 300 python tools/glue_bench.py --reps 10 --out gpurun_out/glue_${TAG}_${cfg%%=*}${cfg##*=}.json > gpurun_out/glue_${TAG}_${cfg%%=*}${cfg##*=}.txt 2>&1
  echo "== [$cfg] exit $?"; cat gpurun_out/glue_${TAG}_${cfg%%=*}${cfg##*=}.txt | tail -45 | grep -v "conv\|gemm"
done
