#!/bin/bash
# float path: all its GPU tests, launch list (batch 32), config-5 bench line
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest -m gpu -q -x -p no:cacheprovider tests/test_gpu_f32.py tests/test_gpu_compat.py tests/test_gpu_ops.py -k "f32 or compat or decode or float" 2>&1 | tail -5
timeout 600 python bench.py --config 2x_fp32 --steps 20 --warmup 5 > gpurun_out/r02_config5_n1.json 2>gpurun_out/r02_config5_n1.err; echo "config5 rc=$?"; tail -n 3 gpurun_out/r02_config5_n1.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/r02_config5_n1.json').read().strip().splitlines()[-1]); print(d['parity'], d['parity_checked'], d['value'], d['ms_per_step'], d['e2e']['value'])"
CODENET_F32_GEMM=tf32x3 ./tools/gpu/r2_f32_launches.sh 2>&1 | tail -14
