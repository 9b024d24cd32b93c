"""GPU box only: one small float-model run (2x COCO geometry, 256x256, batch 2, both GEMM modes), meant to run under compute-sanitizer."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codenet_b200.arch import NetConfig
from codenet_b200.engine_f32 import EngineF32
from codenet_b200.synth import make_raw_state, make_images
cfg = NetConfig(num_classes=80, w2=True)
raw = make_raw_state(cfg, 0)
x = torch.from_numpy(make_images(2, 256, seed=2)).cuda()
for gemm in ("tf32x3", "fp32"):
    eng = EngineF32(cfg, raw, gemm=gemm)
    dets, inds, v = eng.detect(x)
    torch.cuda.synchronize()
    print(gemm, "ok", float(dets[0, 0, 4]), eng.launches)
