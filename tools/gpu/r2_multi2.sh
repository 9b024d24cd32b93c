#!/bin/bash
# 8-GPU box, round-2 final build: scaling runs of config c (N = 8, 4, 2) and config e at N = 8
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { # name N args...
  name=$1; N=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --steps 10 --warmup 3 --no-bilinear "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"; tail -n 2 gpurun_out/$name.err | cut -c1-300
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
    print("$name", "value", j["value"], "e2e", j["e2e"]["value"], "parity", j.get("parity_checked"), "fabric", (j["e2e"].get("fabric") or {}).get("ceiling_images_per_s"), (j["e2e"].get("fabric") or {}).get("h2d_GBps_all_ranks"))
except Exception as e: print("$name parse failed", e)
PY
}
run r02b_config_c_n8 8
run r02b_config_c_n4 4
run r02b_config_c_n2 2
run r02b_config_e_n8 8 --config e
