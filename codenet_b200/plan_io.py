"""On-disk form of a compiled Plan: the "wire format" of a compiled CoDeNet model.

The reference ships `model_last.pth` -- fp32 weights, BatchNorm statistics and QuantAct ranges (lib/models/model.py:35-100) --
and redoes BN folding and weight quantisation on every forward.  A Plan is what is left after that work has been done once:
integer weights, exact requantisation constants, tensor layouts and the op list.  `save_plan` writes it as ONE .npz archive
(plain arrays + a JSON header, no pickle); `load_plan` restores it, and `Engine.from_plan_file` runs it without the checkpoint,
the network definition or torch.load:

    plan = build_plan(cfg, state_dict, 512, 512, "bilinear");  save_plan(plan, "codenet1x_512.cdnplan.npz")
    eng  = Engine.from_plan_file("codenet1x_512.cdnplan.npz", max_batch=256)
"""
import dataclasses
import json

import numpy as np

from .arch import NetConfig
from .plan import Op, Plan, TensorSpec

FORMAT = "codenet_b200.plan/1"


def save_plan(plan: Plan, path: str) -> None:
    arrays = {}
    ops = []
    for i, op in enumerate(plan.ops):
        scal = {}
        for k, v in op.a.items():
            if isinstance(v, np.ndarray):
                arrays["op%d/%s" % (i, k)] = v
            elif isinstance(v, (np.floating, float)):
                scal[k] = {"f": float(v).hex()}                      # hex: exact round trip of fp64 constants
            elif isinstance(v, (np.integer, int, bool)):
                scal[k] = int(v)
            else:
                raise TypeError("op %s: cannot serialise %s of type %s" % (op.name, k, type(v).__name__))
        ops.append({"kind": op.kind, "name": op.name, "scalars": scal})
    cfg = dataclasses.asdict(plan.cfg)
    cfg["heads"] = [list(h) for h in plan.cfg.heads]
    header = {
        "format": FORMAT, "cfg": cfg, "in_H": plan.in_H, "in_W": plan.in_W, "offset_mode": plan.offset_mode, "cat": plan.cat,
        "out_H": plan.out_H, "out_W": plan.out_W, "taps": plan.taps, "ops": ops,
        "tensors": [{"H": t.H, "W": t.W, "C": t.C, "pitch": t.pitch, "half": t.half, "name": t.name,
                     "act": [float(t.act[0]).hex(), float(t.act[1]).hex()]} for t in plan.tensors],
    }
    arrays["header"] = np.frombuffer(json.dumps(header).encode(), dtype=np.uint8)
    np.savez_compressed(path, **arrays)


def load_plan(path: str) -> Plan:
    z = np.load(path, allow_pickle=False)
    header = json.loads(bytes(z["header"]).decode())
    if header.get("format") != FORMAT:
        raise ValueError("%s is not a %s archive" % (path, FORMAT))
    c = header["cfg"]
    c["heads"] = tuple(tuple(h) for h in c["heads"])
    plan = Plan(NetConfig(**c), header["in_H"], header["in_W"], header["offset_mode"])
    for i, t in enumerate(header["tensors"]):
        plan.tensors.append(TensorSpec(i, t["H"], t["W"], t["C"], t["pitch"], t["half"],
                                       (float.fromhex(t["act"][0]), float.fromhex(t["act"][1])), t["name"]))
    for i, o in enumerate(header["ops"]):
        a = {k: (float.fromhex(v["f"]) if isinstance(v, dict) else v) for k, v in o["scalars"].items()}
        pre = "op%d/" % i
        for k in z.files:
            if k.startswith(pre):
                a[k[len(pre):]] = z[k]
        plan.ops.append(Op(o["kind"], o["name"], a))
    plan.cat, plan.out_H, plan.out_W = header["cat"], header["out_H"], header["out_W"]
    plan.taps = {k: int(v) for k, v in header["taps"].items()}
    return plan
