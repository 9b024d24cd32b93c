"""GPU box only: per-phase cycle accounting of unit_fused_kernel (block 0, warp 7) for one stage-2 and one stage-3 unit."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codenet_b200 import _lib
from codenet_b200.arch import NetConfig
from codenet_b200.engine import Engine
from codenet_b200.synth import make_quant_state, make_images
L = _lib.load()
L.cdn_debug_read_cycles.argtypes = [C.c_void_p, C.c_int]
cfg = NetConfig(num_classes=20)
calib = np.load("tests/golden/codenet1x_calib.npz")
st = make_quant_state(cfg, calib, "round", 512)
B = 256
extra = int(os.environ.get("EXTRA", "0"))
L.cdn_set_debug_flags(2 | extra)                      # eager launches (no graph)
eng = Engine.from_state_dict(cfg, st, 512, 512, B, offset_mode="round")
x = torch.from_numpy(np.concatenate([make_images(16, 512, seed=100)] * (B // 16))).cuda()
eng.run(x, maps=False); torch.cuda.synchronize()
buf = (C.c_ulonglong * 16)()
L.cdn_debug_read_cycles(buf, 1)
L.cdn_set_debug_flags(2 | (1 << 21) | extra)
eng.run(x, maps=False); torch.cuda.synchronize()
L.cdn_debug_read_cycles(buf, 1)
v = list(buf)
names = ["wait A1+G1", "E1", "sync1", "S", "sync2", "wait pass+G2", "E2", "sync3"]
print("all 10 fused units, block 0 warp 7: %d cycles total (%.1f us), %d tiles -> %.0f cycles per tile" % (v[8], v[8] / 1965.0, v[9], v[8] / max(v[9], 1)))
print("   " + "  ".join("%s %.1f%%" % (n, 100.0 * c / max(v[8], 1)) for n, c in zip(names, v)))
print("   thread 0 of block 0, cycles per tile: wait A1 %.0f, wait G1 %.0f, wait pass %.0f, wait G2 %.0f" % tuple(c / max(v[9], 1) for c in v[10:14]))
