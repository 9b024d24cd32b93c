#!/bin/bash
# unit_fused.cu: per-phase cycles (block 0) + ncu --set full of the first four launches (3 x stage 2, 1 x stage 3)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/uf_phase_cycles.py 2>&1 | tail -n 4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:unit_fused -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-2} -f -o gpurun_out/prof_full_unit_fused \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-bilinear > gpurun_out/ncu_unit_fused.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_unit_fused.log
