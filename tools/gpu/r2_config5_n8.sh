#!/bin/bash
# 8-GPU box: BASELINE config 5 (2x fp32 COCO, batch 1024 = 128 per GPU) sharded over 8 GPUs, then 2 GPUs
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for N in 8 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --steps 10 --warmup 3 --config 2x_fp32 > gpurun_out/r02_config5_n$N.json 2> gpurun_out/r02_config5_n$N.err
  echo "n$N rc=$?"; tail -n 2 gpurun_out/r02_config5_n$N.err | cut -c1-300
  python -c "
import json; d=json.loads(open('gpurun_out/r02_config5_n$N.json').read().strip().splitlines()[-1]); print('N=$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_checked'])"
done
