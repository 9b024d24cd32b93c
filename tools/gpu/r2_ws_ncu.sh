#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:unit_fused_ws -s 7 -c 1 -f -o gpurun_out/prof_full_unit_ws \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-bilinear --parity-images 1 > gpurun_out/ncu_unit_ws.log 2>&1
echo "ncu rc=$?"
