"""GPU box only: device throughput of the w2 + stride-2/MaxPool configuration at 512x512 (BASELINE config 4 geometry,
per-GPU share) through the same engine.  The calibration archive of this configuration was made at 256x256
(tests/golden/codenet_w2mp_calib.npz); its frozen ranges are reused at 512x512 -- throughput does not depend on them, and
parity of this configuration is tested at 256x256 (tests/test_gpu_engine.py::test_engine_w2_maxpool_matches_reference_vectors)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codenet_b200.arch import NetConfig  # noqa: E402
from codenet_b200.engine import Engine  # noqa: E402
from codenet_b200.synth import make_quant_state, make_images  # noqa: E402

B = int(os.environ.get("B", "256"))
cfg = NetConfig(num_classes=20, w2=True, maxpool=True)
calib = dict(np.load(os.path.join(ROOT, "tests", "golden", "codenet_w2mp_calib.npz")))
st = make_quant_state(cfg, calib, "round", 256)
eng = Engine.from_state_dict(cfg, st, 512, 512, B, offset_mode="round")
x = torch.from_numpy(np.concatenate([make_images(16, 512, seed=100)] * (B // 16))).cuda()
out = {}
for _ in range(3):
    eng.run(x, maps=False, dets=True, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    eng.run(x, maps=False, dets=True, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
pr = eng.profile(x)
fam = {}
for name, kind, t in pr:
    fam[kind] = fam.get(kind, 0.0) + t
row = {"config": "CoDeNet w2 + S2/MaxPool 512x512 W4A8 (BASELINE config 4 geometry), batch %d on one B200" % B,
       "ms_per_step": round(ms, 3), "images_per_s": round(B / ms * 1e3, 1), "requant": dict(zip(("int", "guarded"), eng.requant_stats)),
       "family_ms": {k: round(v, 3) for k, v in fam.items()}}
print(json.dumps(row))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(row, open("gpurun_out/bench_w2.json", "w"), indent=1)
