#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest -m gpu -q --timeout 900 -p no:cacheprovider tests > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/t_gpu.log
./tools/gpu/sanitize.sh
