#!/bin/bash
# ncu --set full of selected kernels: usage scripts_gpu_ncu.sh <regex> <skip> <count> [batch]
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
K=$1; S=$2; C=$3; B=${4:-64}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o gpurun_out/prof_$K \
   python bench.py --steps 1 --warmup 3 --batch $B --no-cpu > gpurun_out/ncu_$K.log 2>&1
echo "ncu $K rc=$?"; tail -n 3 gpurun_out/ncu_$K.log
