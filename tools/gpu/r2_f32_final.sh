#!/bin/bash
# float path, final records: compute-sanitizer over a small run, ncu --set full of the tensor-core GEMM on one layer
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_f32.py > gpurun_out/sanitize_f32_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok " gpurun_out/sanitize_f32_$tool.log | tail -4
done
./tools/gpu/r2_tf32_ncu.sh layer4
./tools/gpu/r2_tf32_ncu.sh l2.pw
