#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest -m gpu -q -x --timeout 300 -p no:cacheprovider tests/test_gpu_ops.py -k "stem" tests/test_gpu_engine.py -k "stem or reference_vectors or u8 or uint8" > gpurun_out/t_stem.log 2>&1
rc=$?; echo "stem tests rc=$rc"; tail -n 12 gpurun_out/t_stem.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-bilinear --dump-ops gpurun_out/ops_uf.json > gpurun_out/bench_uf.log 2>gpurun_out/bench_uf.err
echo "bench rc=$?"; tail -n 3 gpurun_out/bench_uf.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-bilinear --config e > gpurun_out/bench_e.log 2>gpurun_out/bench_e.err
echo "bench e rc=$?"; tail -n 3 gpurun_out/bench_e.err
