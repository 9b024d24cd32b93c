#!/bin/bash
# final round-2 build: isolated deformable-layer sweep (config 2), config e and the bilinear network on one GPU
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python tools/deform_sweep.py > gpurun_out/deform_sweep.log 2>&1; echo "sweep rc=$?"; tail -n 3 gpurun_out/deform_sweep.log | cut -c1-300
timeout 500 python bench.py --steps 10 --warmup 3 --config e --no-bilinear > gpurun_out/r02_config_e_n1.json 2>gpurun_out/r02_config_e_n1.err; echo "config e rc=$?"
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu --offset-mode bilinear > gpurun_out/r02_bilinear_n1.json 2>gpurun_out/r02_bilinear_n1.err; echo "bilinear rc=$?"
