#!/bin/bash
# ncu --set full with source counters for the lazy-upsample strip kernels (second pass of tools/hot_path_once.py)
TAG=${1:-lz}
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"lazy_strip_kernel" -s 4 -c 4 -o gpurun_out/lazy_$TAG python tools/hot_path_once.py --tier-b > gpurun_out/ncu_lazy_$TAG.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_lazy_$TAG.log; ls -la gpurun_out/lazy_$TAG.ncu-rep
