"""GPU box only: per-phase cycle accounting of the pointwise GEMM kernel (block 0) for a few layers."""
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
L.cdn_set_debug_flags(2)                      # eager launches (no graph)
eng = Engine.from_state_dict(cfg, st, 512, 512, B, offset_mode="round")
x = torch.from_numpy(np.concatenate([make_images(16, 512, seed=100)] * (B // 16))).cuda()
eng.run(x, maps=False); torch.cuda.synchronize()
names = ["prod EMPTY", "prod PEMPTY", "mma TEMPTY", "mma FULL", "epi TFULL", "epi PFULL", "epi SEMPTY", "epi chunks", "epi fence+arrive", "st SFULL", "st issue+read", "kernel total", "launches"]
buf = (C.c_ulonglong * 16)()
L.cdn_debug_read_cycles(buf, 1)
pw_ops = [op for op in eng.plan.ops if op.kind == "pw"]
extra = int(os.environ.get("EXTRA", "0"))
for sel in [int(v) for v in os.environ.get("SEL", "0").split(",")]:
    L.cdn_set_debug_flags(2 | 16 | extra | (sel << 8))
    eng.run(x, maps=False); torch.cuda.synchronize()
    L.cdn_debug_read_cycles(buf, 1)
    v = list(buf)
    title = "all pw launches" if sel == 0 else pw_ops[sel - 1].name
    print("%s  (extra flags %d): kernel %d cycles = %.1f us, launches %d" % (title, extra, v[11], v[11] / 1965.0, v[12]))
    print("   " + "  ".join("%s %.0f%%" % (n, 100.0 * c / max(v[11], 1)) for n, c in zip(names[:11], v)))
