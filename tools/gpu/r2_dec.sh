#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest -m gpu -q -x --timeout 300 -p no:cacheprovider tests -k "decode or detector or engine_matches or batch_256" > gpurun_out/t_dec.log 2>&1
echo "decode tests rc=$?"; tail -n 6 gpurun_out/t_dec.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-bilinear --dump-ops gpurun_out/ops_uf.json > gpurun_out/bench_uf.log 2>gpurun_out/bench_uf.err
echo "bench rc=$?"; tail -n 3 gpurun_out/bench_uf.err
