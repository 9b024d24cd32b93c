"""Deterministic synthetic CoDeNet weights and images (there is no network for checkpoints or datasets).

The weights are drawn from a numpy PCG64 stream so that the build container (where the unmodified reference
turns them into golden vectors) and the GPU box (where only this file exists) see identical tensors; a digest
of the generated tensors is stored beside every golden vector.  The recipe follows SURVEY.md 8(d)/F7: stock
init lets the signal die (F7), so convolutions get a variance-preserving normal init, the offset-scale conv
gets N(0,(2/sqrt(C))^2) so the dilation scalar s spans its bound, and the heatmap output conv N(0,0.05^2) so the
sigmoid stays unsaturated and the top-K is tie-free.  BatchNorm running statistics and QuantAct ranges come
from a calibration pass of the reference itself and are shipped in tests/golden/ (they are a few KB).
"""
import hashlib
from typing import Dict

import numpy as np

from .arch import NetConfig, build_graph, raw_param_shapes


def make_raw_state(cfg: NetConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """fp32 tensors for every key of the raw (pre-quantisation) state dict, BN stats at their defaults."""
    g = build_graph(cfg)
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    kinds = {}
    for c in g.all_convs():
        kinds[c.raw_conv + ".weight"] = c
    for key, shape in raw_param_shapes(g).items():
        if key.endswith("running_mean"):
            v = np.zeros(shape)
        elif key.endswith("running_var"):
            v = np.ones(shape)
        elif key in kinds:
            c = kinds[key]
            fan_in = shape[1] * shape[2] * shape[3]
            if c.kind == "scale":
                std = 2.0 / np.sqrt(c.cin)
            elif c.kind == "head_out":
                std = 0.05 if c.name.startswith("hm") else 0.1
            else:
                std = np.sqrt(2.0 / fan_in)
            v = rng.standard_normal(shape) * std
        elif key.endswith(".bias") and key[:-5] in {c.raw_conv for c in g.all_convs()}:
            c = [c for c in g.all_convs() if c.raw_conv == key[:-5]][0]
            if c.kind == "scale":
                v = np.ones(shape)                      # dcn_deform_conv.py:297-302: bias 1 => s = 1 at rest
            elif c.name.startswith("hm"):
                v = np.full(shape, -2.19)               # shufflenetv2_dcn.py:259-262
            else:
                v = rng.standard_normal(shape) * 0.5 + (4.0 if c.name.startswith("wh") else 0.5)
        elif key.endswith(".weight"):                   # BN gamma
            v = rng.uniform(0.6, 1.4, shape)
        else:                                           # BN beta
            v = rng.standard_normal(shape) * 0.2
        out[key] = np.ascontiguousarray(v, dtype=np.float32)
    return out


def state_digest(state: Dict[str, np.ndarray]) -> str:
    h = hashlib.sha256()
    for k in sorted(state):
        h.update(k.encode())
        h.update(np.ascontiguousarray(state[k]).tobytes())
    return h.hexdigest()


def make_images(batch: int, res: int, seed: int = 2, clamp: float = 2.5) -> np.ndarray:
    """Normalised fp32 images [B,3,R,R] (what pre_process hands to the net, base_detector.py:66-70): smooth
    blobs + noise, clamped so a calibration on one batch covers another."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.standard_normal((batch, 3, res, res))
    # low-frequency content so deformable layers and the heatmap see structure, not white noise
    lo = rng.standard_normal((batch, 3, res // 16, res // 16))
    lo = np.repeat(np.repeat(lo, 16, axis=2), 16, axis=3)
    x = 0.6 * x + 0.9 * lo
    return np.clip(x, -clamp, clamp).astype(np.float32)


def make_quant_state(cfg: NetConfig, calib: Dict[str, np.ndarray], mode: str = "round", res: int = 256,
                     seed: int = 0) -> Dict[str, np.ndarray]:
    """State dict in the reference's QUANTISED key space (what `quantize_shufflenetv2_dcn` + `load_model`
    leave in memory, base_detector.py:29-36): synthetic weights + calibrated BN statistics + frozen QuantAct
    ranges taken from a calibration archive (tests/golden/codenet1x_calib.npz)."""
    from .arch import act_keys, raw_to_quant_key
    g = build_graph(cfg)
    raw = make_raw_state(cfg, seed)
    if "digest" in calib and str(calib["digest"]) != state_digest(raw):
        raise RuntimeError("synthetic weights differ from the ones the calibration archive was made with")
    for k in list(raw):
        if "bn/" + k in calib:
            raw[k] = np.asarray(calib["bn/" + k], dtype=np.float32)
    r2q = raw_to_quant_key(g)
    st = {r2q[k]: v for k, v in raw.items()}
    for lbl, p in act_keys(g).items():
        lo, hi = calib["ranges_%s_%d/%s" % (mode, res, lbl)]
        st[p + ".x_min"] = np.array([lo], dtype=np.float32)
        st[p + ".x_max"] = np.array([hi], dtype=np.float32)
    return st
