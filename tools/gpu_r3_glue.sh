#!/bin/bash
# round-2 (session 2) check of the rewritten layout kernels: parity tests that cover them, then the trunk-less per-kernel
# bench with the new forms (default), the round-1 forms (ASN_GLUE=0) and without the lazy-upsample prefetch.
# usage: tools/gpu_r3_glue.sh TAG [test files...]
TAG=${1:-g1}; shift
TESTS=${@:-tests/test_gpu_aspp.py tests/test_gpu_fcd.py tests/test_gpu_lazy.py tests/test_gpu_step.py}
mkdir -p gpurun_out
timeout 600 python -m pytest $TESTS -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/glue_test_$TAG.log 2>&1
echo "pytest exit $?"; grep -E "^E  |passed|failed|FAILED|Error" gpurun_out/glue_test_$TAG.log | cut -c1-300 | head -30
for cfg in ${CFGS:-A=1 ASN_GLUE=0 ASN_LAZY_CHAINS=1}; do
  env $cfg timeout 300 python tools/glue_bench.py --reps 10 --out gpurun_out/glue_${TAG}_${cfg%%=*}${cfg##*=}.json > gpurun_out/glue_${TAG}_${cfg%%=*}${cfg##*=}.txt 2>&1
  echo "== [$cfg] exit $?"; cat gpurun_out/glue_${TAG}_${cfg%%=*}${cfg##*=}.txt | tail -45 | grep -v "conv\|gemm"
done
