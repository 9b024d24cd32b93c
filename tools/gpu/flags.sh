#!/bin/bash
# bench under several CODENET_DEBUG_FLAGS values: usage flags.sh 0 32 64
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
for f in "$@"; do
  CODENET_DEBUG_FLAGS=$f python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_flags$f.json 2>/dev/null
  python - $f <<'PY'
import json, sys
j=json.loads(open('gpurun_out/bench_flags%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print("flags", sys.argv[1], j["value"], {k:v["ms"] for k,v in j["roofline"]["families"].items()})
PY
done
