#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest -m gpu -q -x --timeout 600 -p no:cacheprovider tests/test_gpu_compat.py > gpurun_out/t_compat.log 2>&1
echo "pytest compat rc=$?"; tail -n 15 gpurun_out/t_compat.log
timeout 600 python bench.py --config 2x_fp32 --steps 5 --warmup 3 > gpurun_out/bench_f32.log 2>gpurun_out/bench_f32.err; echo "f32 rc=$?"; tail -c 900 gpurun_out/bench_f32.log; tail -n 5 gpurun_out/bench_f32.err
