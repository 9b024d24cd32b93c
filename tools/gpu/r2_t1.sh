#!/bin/bash
# round 2, trip 1: full GPU test-suite + the default bench line (parity check, pipelined e2e, fabric, bilinear leg, reference CPU arm)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest -m gpu -q -x --timeout 900 -p no:cacheprovider tests ${PYTEST_ARGS} > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 12 gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/ops.json > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err
echo "ref rc=$?"; tail -c 1500 gpurun_out/bench_ref.log | cut -c1-400; tail -n 3 gpurun_out/bench_ref.err
