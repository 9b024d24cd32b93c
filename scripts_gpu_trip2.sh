#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest -m gpu -q --timeout 300 -p no:cacheprovider tests/test_gpu_ops.py -k "decode" > gpurun_out/t1_decode.log 2>&1
echo "decode rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --dump-ops gpurun_out/ops_r1_v1.json > gpurun_out/bench_v1.log 2>&1
echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v1.csv python bench.py --steps 1 --warmup 3 --batch 64 > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"
tail -n 3 gpurun_out/t1_decode.log
