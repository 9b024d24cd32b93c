"""GPU box only: one small engine run (256x256, batch 3) + stand-alone ops, meant to run under compute-sanitizer."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codenet_b200.arch import NetConfig
from codenet_b200.engine import Engine
from codenet_b200.synth import make_quant_state, make_images
cfg = NetConfig(num_classes=20)
calib = np.load(os.path.join(ROOT, "tests/golden/codenet1x_calib.npz"))
for mode in ("round", "bilinear"):
    st = make_quant_state(cfg, calib, mode, 256)
    eng = Engine.from_state_dict(cfg, st, 256, 256, 3, offset_mode=mode)
    eng.set_option("use_graph", 0)
    x = torch.from_numpy(make_images(3, 256, seed=2)).cuda()
    out = eng.run(x)
    torch.cuda.synchronize()
    print(mode, "ok", float(out["dets"][0, 0, 4]))
    eng.close()
