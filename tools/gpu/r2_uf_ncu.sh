#!/bin/bash
# fused unit kernels: ncu --set full of selected launches of the first warm-up step (order per step: s2, 3 x stage 2, 7 x stage 3)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:unit_ -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-5} -f -o gpurun_out/prof_full_unit_fused \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-bilinear > gpurun_out/ncu_unit_fused.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_unit_fused.log | cut -c1-300
