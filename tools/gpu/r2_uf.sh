#!/bin/bash
# GPU trip for unit_fused.cu: its parity tests under a short timeout, the engine tests, bench with and without the unit fusion
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest -m gpu -q -x --timeout 300 -p no:cacheprovider tests/test_gpu_engine.py -k "units_fused or reference_vectors_256" > gpurun_out/t_uf.log 2>&1
rc=$?; echo "uf tests rc=$rc"; tail -n 25 gpurun_out/t_uf.log
if [ $rc -ne 0 ]; then exit $rc; fi
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-bilinear --dump-ops gpurun_out/ops_uf.json > gpurun_out/bench_uf.log 2>gpurun_out/bench_uf.err
echo "bench fused rc=$?"; tail -c 3000 gpurun_out/bench_uf.log; tail -n 5 gpurun_out/bench_uf.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-bilinear --no-fuse-units > gpurun_out/bench_nouf.log 2>gpurun_out/bench_nouf.err
echo "bench unfused rc=$?"; tail -c 1500 gpurun_out/bench_nouf.log; tail -n 5 gpurun_out/bench_nouf.err
${EXTRA_CMD:-true}
