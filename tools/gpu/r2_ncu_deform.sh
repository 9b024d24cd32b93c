#!/bin/bash
# ncu --set full of the three deform_tile launches of one step (batch 256), plus the full GPU test-suite
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deform_tile -s 9 -c 3 -f -o gpurun_out/prof_deform_tile \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-bilinear > gpurun_out/ncu_deform_tile.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_deform_tile.log | cut -c1-300
timeout 1500 python -m pytest -m gpu -q -x --timeout 900 -p no:cacheprovider tests > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 12 gpurun_out/t_gpu.log
