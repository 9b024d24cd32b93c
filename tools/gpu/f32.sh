#!/bin/bash
# float (fp32) path: parity tests + throughput (config 5 geometry on one GPU)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest -m gpu -q -x -p no:cacheprovider tests/test_gpu_f32.py tests/test_gpu_compat.py 2>&1 | tail -3
timeout 600 python tools/bench_f32.py 2>&1 | tail -2 | tee gpurun_out/bench_f32.log
