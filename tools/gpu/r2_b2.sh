#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest -m gpu -q -x --timeout 600 -p no:cacheprovider tests/test_gpu_modules.py > gpurun_out/t_b2.log 2>&1
echo "pytest b2 rc=$?"; tail -n 30 gpurun_out/t_b2.log
timeout 900 python tools/deform_sweep.py > gpurun_out/deform_sweep.log 2>&1; echo "sweep rc=$?"; tail -n 3 gpurun_out/deform_sweep.log | cut -c1-300
