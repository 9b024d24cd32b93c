#!/bin/bash
# GPU trip: parity tests, bench, launch list, ncu --set full of the main kernel families
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest -m gpu -q --timeout 600 -p no:cacheprovider tests > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/ops.json > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench.log
for k in pw_gemm_tc dw3x3 deform_dw ctdet_decode; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 4 -f -o gpurun_out/prof_$k \
     python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out
