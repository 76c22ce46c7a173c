#!/bin/bash
# first bring-up run on the B200 box: each suite under its own timeout so a hang in one
# does not hide the others.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
ls /root/reference > gpurun_out/ref_ls.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" > gpurun_out/cpu.txt 2>&1
for t in pointwise umma aspp; do
  timeout 600 python -m pytest tests/test_gpu_$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/test_$t.log 2>&1
  echo "test_$t exit $?" >> gpurun_out/summary.txt
done
timeout 300 python -m pytest tests/test_gpu_aspp.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/test_aspp_all.log 2>&1
echo "test_aspp_all exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 30 gpurun_out/test_umma.log
