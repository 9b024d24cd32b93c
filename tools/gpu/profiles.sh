#!/bin/bash
# Round profiles: (1) launch list with device time + DRAM bytes of every launch of the bench command (batch 256),
# (2) ncu --set full of the dominant kernel (pointwise GEMM) and of the deformable / depthwise kernels.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
   --log-file gpurun_out/launches_full.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
for k in pw_gemm_tc deform_int_v3 dw3x3_v2 heads_fused; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIPK:-3} -c ${COUNTK:-3} -f -o gpurun_out/prof_full_$k \
     python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_$k.log 2>&1
  echo "ncu full $k rc=$?"
done
ls -la gpurun_out | head -30
