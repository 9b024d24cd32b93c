#!/bin/bash
# ncu --set full of the tensor-core float GEMM on one layer shape (argument: layer name of tools/gpu/r2_tf32_layer.py)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=${1:-l2.pw}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pw_tf32x3 -s 3 -c 1 -f -o gpurun_out/prof_full_pw_tf32_$L \
   python tools/gpu/r2_tf32_layer.py $L > gpurun_out/ncu_pw_tf32.log 2>&1
echo "ncu rc=$?"
