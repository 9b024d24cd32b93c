import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codenet_b200.arch import NetConfig
from codenet_b200.synth import make_raw_state, make_images
from codenet_b200.engine_f32 import EngineF32
cfg = NetConfig(num_classes=20)
g = np.load(os.path.join(ROOT, "tests/golden/codenet_float_1x_256.npz"))
raw = make_raw_state(cfg, 0)
for k in g.files:
    if k.startswith("bn/"): raw[k[3:]] = g[k]
eng = EngineF32(cfg, raw)
x = torch.from_numpy(make_images(2, 256, seed=2)[:1].copy()).cuda()
v = eng.forward(x)
for name, key in (("hm", "hm_logit"), ("wh", "wh"), ("reg", "reg")):
    got, ref = v[name].cpu().numpy(), g[key]
    e = np.abs(got - ref)
    print(name, "max", e.max(), "p50", np.percentile(e, 50), "p99", np.percentile(e, 99), "p99.9", np.percentile(e, 99.9), "p99.99", np.percentile(e, 99.99),
          "argmax", np.unravel_index(e.argmax(), e.shape), "scale", np.abs(ref).max())
    big = np.argwhere(e > 20 * np.percentile(e, 99))
    print("  n big", len(big), big[:10].tolist())
