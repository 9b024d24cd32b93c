#!/bin/bash
# unit_fused.cu: ncu --set full of two launches (the third stage-2 unit and the first stage-3 unit of the first warm-up step)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:unit_fused -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-2} -f -o gpurun_out/prof_full_unit_fused \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-bilinear > gpurun_out/ncu_unit_fused.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_unit_fused.log | cut -c1-300
