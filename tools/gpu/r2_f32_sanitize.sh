#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over a small float-model run in both GEMM modes (tools/sanitize_f32.py)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_f32.py > gpurun_out/sanitize_f32_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok " gpurun_out/sanitize_f32_$tool.log | tail -4
done
