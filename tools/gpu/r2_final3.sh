#!/bin/bash
# round-2 closing run on one B200: the whole GPU test suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; tail -n 2 gpurun_out/r02_bench_final.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity_checked'), d['roofline']['frac'], d['gpu_launches'], d['cpu_baseline']['value'] if d.get('cpu_baseline') else None)"
