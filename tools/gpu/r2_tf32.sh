#!/bin/bash
# tensor-core float GEMM: op test, float model parity, config-5 bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest -m gpu -q -x -p no:cacheprovider tests/test_gpu_f32.py 2>&1 | tail -15
timeout 600 python bench.py --config 2x_fp32 --steps 20 --warmup 5 > gpurun_out/r02_config5_n1.json 2>gpurun_out/r02_config5_n1.err; echo "config5 rc=$?"; tail -n 3 gpurun_out/r02_config5_n1.err | cut -c1-300
cut -c1-1200 gpurun_out/r02_config5_n1.json
