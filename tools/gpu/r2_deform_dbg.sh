#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for flags in 0 65536 131072 262144 393216; do
  CODENET_DEBUG_FLAGS=$flags timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-bilinear --parity-images 1 > gpurun_out/bench_d$flags.log 2>gpurun_out/bench_d$flags.err
  echo "bench flags=$flags rc=$?"
  python - <<PY
import json
j=json.loads(open('gpurun_out/bench_d$flags.log').read().strip().splitlines()[-1])
print("flags $flags value", j["value"], "deform", j["deform"]["ms"], [(l["layer"], l["ms"]) for l in j["deform"]["layers"]])
PY
done
timeout 600 python bench.py --config 2x_fp32 --steps 5 --warmup 3 > gpurun_out/bench_f32.log 2>gpurun_out/bench_f32.err; echo "f32 rc=$?"; tail -c 1500 gpurun_out/bench_f32.log; tail -n 5 gpurun_out/bench_f32.err
timeout 600 python bench.py --config e --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_e.log 2>gpurun_out/bench_e.err; echo "e rc=$?"; tail -c 2500 gpurun_out/bench_e.log | cut -c1-2500; tail -n 5 gpurun_out/bench_e.err
