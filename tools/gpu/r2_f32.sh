#!/bin/bash
# float path (BASELINE config 5): its GPU tests, then the bench line at N = 1 (long warm-up, 20 timed steps)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest -m gpu -q -x -p no:cacheprovider tests/test_gpu_f32.py tests/test_gpu_compat.py tests/test_gpu_ops.py -k "f32 or compat or decode" 2>&1 | tail -3
timeout 600 python bench.py --config 2x_fp32 --steps 20 --warmup 5 > gpurun_out/r02_config5_n1.json 2>gpurun_out/r02_config5_n1.err; echo "config5 rc=$?"; tail -n 2 gpurun_out/r02_config5_n1.err | cut -c1-200
tail -c 1500 gpurun_out/r02_config5_n1.json
