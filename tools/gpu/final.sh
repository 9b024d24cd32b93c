#!/bin/bash
# sweep (config 2) + bilinear-mode network + 2-GPU bench when two devices are visible
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python tools/deform_sweep.py > gpurun_out/deform_sweep.log 2>&1; echo "sweep rc=$?"; tail -n 3 gpurun_out/deform_sweep.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --offset-mode bilinear > gpurun_out/bench_bilinear.json 2>gpurun_out/bench_bilinear.err; echo "bilinear rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_bilinear.json').read().strip().splitlines()[-1])
print("bilinear", j["value"], j["ms_per_step"], j["e2e"]["value"], j.get("requant"), {k:v["ms"] for k,v in j["roofline"]["families"].items()})
PY
N=$(python -c "import torch; print(torch.cuda.device_count())")
if [ "$N" -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2>gpurun_out/bench_n2.err; echo "n2 rc=$?"
  tail -c 400 gpurun_out/bench_n2.err; python -c "
import json
j=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print('N=2', j['value'], j['ms_per_step'], j['e2e']['value'], j['n_gpus'])"
fi
