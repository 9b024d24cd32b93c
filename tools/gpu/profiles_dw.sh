#!/bin/bash
# launch list (device time + DRAM bytes of every launch of the bench command) and a --set full capture of the TMA depthwise kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
   --log-file gpurun_out/launches_full.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw3x3_tma -s 57 -c 6 -f -o gpurun_out/prof_full_dw3x3_tma \
   python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_dw3x3_tma.log 2>&1
echo "ncu full dw3x3_tma rc=$?"
