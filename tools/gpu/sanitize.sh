#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over one small engine run in both offset modes
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok " gpurun_out/sanitize_$tool.log | tail -4
done
