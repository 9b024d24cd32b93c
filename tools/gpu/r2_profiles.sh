#!/bin/bash
# Round-2 profiles: (1) launch list with device time + DRAM bytes of every launch of the bench command (batch 256),
# (2) ncu --set full of the dominant kernel (pointwise GEMM), the deformable tile kernels (both offset modes), heads, depthwise.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
   --log-file gpurun_out/launches_full.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-bilinear --parity-images 1 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
for k in pw_gemm_tc deform_tile_int heads_fused dw3x3_tma; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 3 -f -o gpurun_out/prof_full_$k \
     python bench.py --steps 1 --warmup 3 --no-cpu --no-bilinear --parity-images 1 > gpurun_out/ncu_full_$k.log 2>&1
  echo "ncu full $k rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deform_tile_bil -s 3 -c 3 -f -o gpurun_out/prof_full_deform_tile_bil \
   python bench.py --steps 1 --warmup 3 --no-cpu --offset-mode bilinear --parity-images 1 > gpurun_out/ncu_full_deform_tile_bil.log 2>&1
echo "ncu full bil rc=$?"
ls -la gpurun_out | grep -E "prof_full|launches_full"
