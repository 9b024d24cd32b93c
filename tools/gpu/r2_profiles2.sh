#!/bin/bash
# Round-2 final profiles: (1) launch list with device time + DRAM bytes of every launch of the bench command (batch 256),
# (2) ncu --set full of the fused unit kernels (s2 + one stage-2 + one stage-3 unit), the pointwise GEMM, heads, stem, deformable tile.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv \
   --log-file gpurun_out/launches_full.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-bilinear --parity-images 1 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:unit_ -s 11 -c 6 -f -o gpurun_out/prof_full_unit_fused \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-bilinear --parity-images 1 > gpurun_out/ncu_full_unit_fused.log 2>&1
echo "ncu full unit_fused rc=$?"
for k in pw_gemm_tc heads_fused stem_fast deform_tile_int; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 2 -f -o gpurun_out/prof_full_$k \
     python bench.py --steps 1 --warmup 3 --no-cpu --no-bilinear --parity-images 1 > gpurun_out/ncu_full_$k.log 2>&1
  echo "ncu full $k rc=$?"
done
ls -la gpurun_out | grep -E "prof_full|launches_full"
