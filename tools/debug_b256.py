"""GPU box only: run every pw op of config c stand-alone at batch B to find a failing launch."""
import os, sys, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from codenet_b200 import _lib
from codenet_b200.arch import NetConfig
from codenet_b200.plan import build_plan
from codenet_b200.synth import make_quant_state
from gpu_util import run_op
B = int(os.environ.get("B", "256"))
cfg = NetConfig(num_classes=20)
calib = np.load(os.path.join(ROOT, "tests/golden/codenet1x_calib.npz"))
st = make_quant_state(cfg, calib, "round", 512)
plan = build_plan(cfg, st, 512, 512, "round")
rng = np.random.default_rng(0)
for op in plan.ops:
    if op.kind != "pw":
        continue
    a = op.a
    tin = plan.tensors[a["in_t"]]
    T = {tin.id: rng.integers(-128, 128, (B, tin.H, tin.W, tin.pitch), dtype=np.int8)}
    if a["pass_t"] >= 0:
        tp = plan.tensors[a["pass_t"]]
        T[tp.id] = rng.integers(-128, 128, (B, tp.H, tp.W, tp.pitch), dtype=np.int8)
    try:
        out = run_op(plan, op, T)
        torch.cuda.synchronize()
        print(op.name, "ok", out.shape, flush=True)
    except Exception as e:
        print(op.name, "FAILED", str(e)[:200], flush=True)
        break
