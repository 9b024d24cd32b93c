#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest -m gpu -q --timeout 900 -p no:cacheprovider tests ${PYTEST_ARGS} > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 12 gpurun_out/t_gpu.log
