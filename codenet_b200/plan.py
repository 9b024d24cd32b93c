"""Host-side compiler: quantised CoDeNet state dict -> static execution plan for libcodenet_b200.

Input is a state dict in the reference's QUANTISED key space (what `quantize_shufflenetv2_dcn` followed by
`load_model` leaves in memory, lib/detectors/base_detector.py:29-36) with frozen QuantAct ranges.
Everything the reference recomputes on every forward is done once here, in fp64 and in the reference's operation
order so that the integer weights and constants are the ones its fp64 evaluation produces:

  * BN folding                 portable_quantizer/quant_modules.py:364-372
  * per-channel symmetric k-bit weight quantisation       quant_utils.py:76-82, :205-223
  * activation scale / zero point                          quant_utils.py:58-73
  * requantisation constants   M_c = s_out/(sigma_c s_x),  B_c = s_out b'_c - z_out   

and the graph glue (split / cat / channel_shuffle, shufflenetv2_dcn.py:29-34,102-114; nearest upsample :303) is
turned into memory layout decisions:

  * stage tensors use a HALF layout -- logical channels [0,C/2) at bytes [0,C/2), [C/2,C) at [Hp, Hp+C/2),
    Hp = C/2 rounded up to 32 -- so the branch input of a stride-1 unit is one aligned TMA box
  * the last 1x1 conv of every unit writes its channels interleaved with the pass-through half (chunk table)
  * nearest x2 upsampling is virtual: the consumer reads (y>>1, x>>1); 1x1 convs that follow an upsample are
    evaluated before it (pointwise ops commute with nearest upsampling exactly)
  * the three heads run as one 64->192 GEMM, one 192-channel depthwise conv and one block-diagonal 192->(cat+4) GEMM
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .arch import NetConfig, build_graph, act_keys

F = np.float64


def _r(x, m):
    return (x + m - 1) // m * m


# ---- quantisation arithmetic (fp64, reference operation order) -------------------------------------------------
def act_params(lo, hi, bits=8):
    lo, hi = F(lo), F(hi)
    # `n / tensor` in torch is tensor.reciprocal() * n (Tensor.__rtruediv__): two roundings, restated as such
    s = (F(1) / max(hi - lo, F(1e-10))) * F(2 ** bits - 1)
    z = np.rint(s * lo) + F(2 ** (bits - 1))
    return float(s), float(z)


def weight_range(flat, percentile=False):
    """Per-output-channel (w_min, w_max) of the flattened weights [Cout, n] (quant_modules.py:373-395): plain min / max, or with
    --wt-percentile the k-th smallest values at k = ceil(0.1 % n) and ceil(99.9 % n) (1-based, torch.kthvalue); rows of fewer
    than 10 values (every depthwise 3x3 kernel) take 0.95 * min / max instead."""
    lo, hi = flat.min(1), flat.max(1)
    if not percentile:
        return lo, hi
    n = flat.shape[1]
    if n < 10:
        return lo * F(0.95), hi * F(0.95)
    srt = np.sort(flat, axis=1)
    kl, ku = int(np.ceil(n * 0.1 * 0.01)), int(np.ceil(n * 99.9 * 0.01))
    return srt[:, kl - 1], srt[:, ku - 1]


def quant_weight(w, bits, percentile=False):
    """w: [Cout, ...] fp64.  Returns integer weights (int64) and the per-channel scale sigma."""
    flat = w.reshape(w.shape[0], -1)
    lo, hi = weight_range(flat, percentile)
    mag = np.maximum(np.abs(lo), np.abs(hi))
    sigma = (F(1) / np.maximum(mag, F(1e-10))) * F(2 ** (bits - 1) - 1)   # reciprocal * n, as torch evaluates n / tensor
    q = np.rint(sigma.reshape(-1, *([1] * (w.ndim - 1))) * w)
    q = np.clip(q, -(2 ** (bits - 1)), 2 ** (bits - 1) - 1)
    return q.astype(np.int64), sigma


def fold_bn(w, gamma, beta, mean, var, eps=1e-5):
    std = np.sqrt(var + F(eps))
    sf = gamma / std
    return w * sf.reshape(-1, 1, 1, 1), (F(0) - mean) * sf + beta


# ---- plan records -------------------------------------------------------------------------------------------------
@dataclass
class TensorSpec:
    id: int
    H: int                                   # per image, as STORED
    W: int
    C: int                                   # logical channels
    pitch: int                               # bytes per pixel
    half: int = 0                            # 0: dense; else Hp (byte offset of the second half)
    act: tuple = (1.0, 0.0)                  # (s, z) of the producing QuantAct
    name: str = ""

    def phys(self, c):
        c = np.asarray(c)
        if not self.half:
            return c
        h = self.C // 2
        return np.where(c < h, c, self.half + c - h)


@dataclass
class Op:
    kind: str                                # 'stem' | 'dw' | 'deform' | 'pw'
    name: str
    a: dict = field(default_factory=dict)


@dataclass
class Plan:
    cfg: NetConfig
    in_H: int
    in_W: int
    offset_mode: str
    tensors: List[TensorSpec] = field(default_factory=list)
    ops: List[Op] = field(default_factory=list)
    cat: int = 0
    out_H: int = 0
    out_W: int = 0
    taps: Dict[str, int] = field(default_factory=dict)      # label -> tensor id (for tests)

    def add_tensor(self, H, W, C, pitch, half=0, act=(1.0, 0.0), name=""):
        t = TensorSpec(len(self.tensors), H, W, C, pitch, half, act, name)
        self.tensors.append(t)
        return t


def _requant(s_x, sigma, b, out_act, relu):
    s_o, z_o = out_act
    M = F(s_o) / (sigma * F(s_x))
    B = F(s_o) * b - F(z_o)
    lo = max(-128, int(-z_o)) if relu else -128
    return M, B, lo


def _zx(act):
    z = int(act[1])
    if not (-127 <= z <= 128):
        raise ValueError("activation range does not contain 0 (zero point %d): real zero is not on the int8 grid" % z)
    return z


class PlanBuilder:
    def __init__(self, cfg: NetConfig, state: Dict[str, np.ndarray], in_H: int, in_W: int, offset_mode="round"):
        if offset_mode not in ("round", "bilinear"):
            raise ValueError("offset_mode must be 'round' or 'bilinear'")
        if in_H % 32 or in_W % 32:
            raise ValueError("input size must be a multiple of 32 (got %dx%d)" % (in_H, in_W))
        self.cfg, self.st, self.g = cfg, state, build_graph(cfg)
        self.ak = act_keys(self.g)
        self.plan = Plan(cfg, in_H, in_W, offset_mode)

    # -- parameters ------------------------------------------------------------------------------------------
    def act(self, label):
        p = self.ak[label]
        return act_params(np.asarray(self.st[p + ".x_min"]).reshape(-1)[0],
                          np.asarray(self.st[p + ".x_max"]).reshape(-1)[0], self.cfg.a_bit)

    def weights(self, c):
        w = np.asarray(self.st[c.q_conv + ".weight"]).astype(F)
        if c.q_bn:
            w, b = fold_bn(w, *(np.asarray(self.st[c.q_bn + "." + f]).astype(F)
                                for f in ("weight", "bias", "running_mean", "running_var")))
        elif c.has_bias:
            b = np.asarray(self.st[c.q_conv + ".bias"]).astype(F)
        else:
            b = np.zeros(c.cout, F)
        wq, sigma = quant_weight(w, c.w_bit, getattr(self.cfg, "wt_percentile", False))
        return wq, sigma, b

    # -- op emitters -----------------------------------------------------------------------------------------
    def emit_dw(self, c, tin: TensorSpec, out_act, relu, stride, in_shift, name, extra=None):
        """Depthwise conv keeps every channel at its physical position."""
        P = self.plan
        wq, sigma, b = self.weights(c)
        M, B, lo = _requant(tin.act[0], sigma, b, out_act, relu)
        H, W = tin.H << in_shift, tin.W << in_shift
        tout = P.add_tensor((H - 1) // stride + 1, (W - 1) // stride + 1, tin.C, tin.pitch, tin.half, out_act, name)
        ph = tin.phys(np.arange(tin.C))
        wp = np.zeros((tin.pitch, 9), np.int8)
        Mp, Bp = np.zeros(tin.pitch, F), np.zeros(tin.pitch, F)
        wp[ph] = wq.reshape(-1, 9)
        Mp[ph], Bp[ph] = M, B
        a = dict(in_t=tin.id, out_t=tout.id, in_shift=in_shift, stride=stride, wq=wp, C=tin.pitch, zx=_zx(tin.act),
                 M=Mp, B=Bp, lo=lo)
        if extra:
            a.update(extra)
        P.ops.append(Op("deform" if extra else "dw", name, a))
        return tout

    def emit_pw(self, convs, tin: TensorSpec, k_off, K_phys, out_acts, relu, name, *, out=None, interleave_with=None,
                f32=False, kmap=None):
        """1x1 conv(s) reading bytes [k_off, k_off+K_phys) of tin.

        convs: list of (ConvSpec, input-channel physical positions relative to k_off) fused along N; every conv
        sees the same input tensor.  out: existing output TensorSpec (interleave case) or None (new dense tensor).
        """
        P = self.plan
        rows, Ms, Bs, Mf, bf = [], [], [], [], []
        lo = None
        for (c, kpos), oa in zip(convs, out_acts):
            wq, sigma, b = self.weights(c)
            wq = wq.reshape(c.cout, c.cin)
            full = np.zeros((c.cout, K_phys), np.int8)
            full[:, kpos] = wq
            rows.append(full)
            if f32:
                Mf.append(F(1) / (sigma * F(tin.act[0]) if kmap is None else sigma * F(kmap[c.name])))
                bf.append(b)
            else:
                s_x = tin.act[0] if kmap is None else kmap[c.name]
                M, B, l = _requant(s_x, sigma, b, oa, relu)
                Ms.append(M); Bs.append(B)
                if lo is not None and l != lo:
                    raise ValueError("fused 1x1 convs need equal output zero points")
                lo = l
        wcat = np.concatenate(rows, 0)
        n_real = wcat.shape[0]
        a = dict(in_t=tin.id, k_off=k_off, K=K_phys, zx=_zx(tin.act), pass_t=-1)
        if f32:
            N = _r(n_real, 16)
            w = np.zeros((N, K_phys), np.int8); w[:n_real] = wcat
            a.update(N=N, wq=w, n_f32=n_real, Mf=np.concatenate(Mf), bf=np.concatenate(bf), out_t=-1)
            P.ops.append(Op("pw", name, a))
            return None
        Mc, Bc = np.concatenate(Ms), np.concatenate(Bs)
        chunks = []
        if interleave_with is None:
            # dense output: GEMM column n -> byte n
            N = _r(n_real, 16)
            pitch = _r(n_real, 32)
            tout = P.add_tensor(tin.H, tin.W, n_real, pitch, 0, out_acts[0], name)
            col = np.arange(n_real)
            for j in range(pitch // 16):
                cnt = int(np.clip(n_real - 16 * j, 0, 16))
                chunks.append((16 * j if cnt else 0, cnt, -1, 16 * j))
        else:
            # channel shuffle: logical out channel 2k = pass[k], 2k+1 = new[k]; out tensor has the HALF layout
            tpass, pass_off0 = interleave_with
            half = n_real
            tout = out
            assert tout.C == 2 * half and half % 2 == 0 and tout.half
            per_group = half // 2
            Gp = _r(per_group, 8)
            if 2 * Gp > 256:                 # two N tiles: each group starts its own tile
                Gp = _r(per_group, 256)
            N = _r(2 * Gp, 16)
            col = np.where(np.arange(half) < per_group, np.arange(half), Gp + np.arange(half) - per_group)
            for g in range(2):
                for j in range(tout.half // 16):
                    cnt = int(np.clip(per_group - 8 * j, 0, 8))
                    if cnt:
                        chunks.append((g * Gp + 8 * j, cnt, pass_off0 + g * per_group + 8 * j, g * tout.half + 16 * j))
                    else:
                        chunks.append((0, 0, -1, g * tout.half + 16 * j))
            a["pass_t"] = tpass.id
        w = np.zeros((N, K_phys), np.int8)
        Mn, Bn = np.zeros(N, F), np.zeros(N, F)
        w[col] = wcat
        Mn[col], Bn[col] = Mc, Bc
        a.update(N=N, wq=w, M=Mn, B=Bn, lo=lo, chunks=np.array(chunks, np.int16).reshape(-1, 4), n_f32=0, out_t=tout.id)
        P.ops.append(Op("pw", name, a))
        return tout

    # -- the network ------------------------------------------------------------------------------------------
    def build(self) -> Plan:
        P, g, cfg = self.plan, self.g, self.cfg
        # stem (fp32 image in, 8-bit weights) -------------------------------------------------------------
        wq, sigma, b = self.weights(g.stem)
        a0 = self.act("stem")
        s0, z0 = a0
        H0, W0 = P.in_H // 4, P.in_W // 4
        t = P.add_tensor(H0, W0, g.stem.cout, 32, 0, a0, "stem")
        P.ops.append(Op("stem", "layer0", dict(out_t=t.id, H=P.in_H, W=P.in_W, stride=g.stem.stride,
                                               pool=1 if cfg.maxpool else 0, wq=wq.astype(np.int8), C=g.stem.cout,
                                               M=F(s0) / sigma, B=F(s0) * b - F(z0), lo=max(-128, int(-z0)))))
        P.taps["stem"] = t.id
        x = t
        # ShuffleNetV2 stages ----------------------------------------------------------------------------------
        for u in g.units:
            r = "layer%d.%d." % (u["stage"], u["unit"])
            cv = u["convs"]
            shared = self.act("layer%d.shared" % u["stage"])
            a1, a2 = self.act(r + "act1"), self.act(r + "act2")
            half = u["oup"] // 2
            Hp = _r(half, 32)
            if u["stride"] == 2:
                kpos = x.phys(np.arange(x.C))
                a4 = self.act(r + "act4")
                d4 = self.emit_dw(cv["dw4"], x, a4, False, 2, 0, r + "dw4")
                P.taps[r + "act4"] = d4.id
                x1 = self.emit_pw([(cv["pw5"], kpos)], d4, 0, x.pitch, [shared], True, r + "pw5")
                c1 = self.emit_pw([(cv["pw1"], kpos)], x, 0, x.pitch, [a1], True, r + "pw1")
                P.taps[r + "act1"] = c1.id
                d2 = self.emit_dw(cv["dw2"], c1, a2, False, 2, 0, r + "dw2")
                P.taps[r + "act2"] = d2.id
                out = P.add_tensor(d2.H, d2.W, u["oup"], 2 * Hp, Hp, shared, r + "out")
                self.emit_pw([(cv["pw3"], np.arange(half))], d2, 0, d2.pitch, [shared], True, r + "pw3",
                             out=out, interleave_with=(x1, 0))
            else:
                assert x.half == Hp and x.C == u["oup"]
                c1 = self.emit_pw([(cv["pw1"], np.arange(half))], x, Hp, Hp, [a1], True, r + "pw1")
                P.taps[r + "act1"] = c1.id
                d2 = self.emit_dw(cv["dw2"], c1, a2, False, 1, 0, r + "dw2")
                P.taps[r + "act2"] = d2.id
                out = P.add_tensor(x.H, x.W, u["oup"], 2 * Hp, Hp, shared, r + "out")
                self.emit_pw([(cv["pw3"], np.arange(half))], d2, 0, d2.pitch, [shared], True, r + "pw3",
                             out=out, interleave_with=(x, 0))
            x = out
            P.taps["layer%d.out" % u["stage"]] = x.id
        # layer4 ---------------------------------------------------------------------------------------------------
        a4 = self.act("layer4")
        x = self.emit_pw([(g.layer4, x.phys(np.arange(x.C)))], x, 0, x.pitch, [a4], True, "layer4")
        P.taps["layer4"] = x.id
        # up path: deformable depthwise + 1x1 (+BN, ReLU) ; the x2 upsample is read virtually by the next op ---------
        shift = 0
        for up in g.ups:
            i = up["idx"]
            wq_s, sigma_s, b_s = self.weights(up["scale"])
            ss, zs = self.act("up%d.s" % i)
            ws = np.zeros(x.pitch, np.int8)
            ws[:x.C] = wq_s.reshape(-1)
            extra = dict(ws=ws, Ms=float(F(1) / (sigma_s[0] * F(x.act[0]))), bs=float(b_s[0]), ss=ss, zs=zs,
                         bound=cfg.offset_bound, mode=0 if P.offset_mode == "round" else 1)
            ad = self.act("up%d.deform" % i)
            d = self.emit_dw(up["deform"], x, ad, False, 1, shift, "up%d.deform" % i, extra=extra)
            P.taps["up%d.deform" % i] = d.id
            ao = self.act("up%d.out" % i)
            x = self.emit_pw([(up["channel"], np.arange(d.C))], d, 0, d.pitch, [ao], True, "up%d.channel" % i)
            P.taps["up%d.out" % i] = x.id
            shift = 1
        # heads: pw1 (x3 fused, evaluated BEFORE the last upsample), dw2 through the virtual upsample, fused out conv
        hs = g.heads
        a1s = [self.act(h["name"] + ".act1") for h in hs]
        a3s = [self.act(h["name"] + ".act3") for h in hs]
        if len({int(a[1]) for a in a1s}) != 1 or len({int(a[1]) for a in a3s}) != 1:
            raise ValueError("fused heads need equal zero points in the head activations")
        hp = self.emit_pw([(h["pw1"], np.arange(64)) for h in hs], x, 0, x.pitch, a1s, True, "heads.pw1")
        hp.act = a1s[0]
        P.taps["heads.act1"] = hp.id
        # fused depthwise over 3x64 channels, each third with its own scales
        wqs, Ms, Bs = [], [], []
        for h, a1, a3 in zip(hs, a1s, a3s):
            wq, sigma, b = self.weights(h["dw2"])
            M, B, lo = _requant(a1[0], sigma, b, a3, True)
            wqs.append(wq.reshape(-1, 9)); Ms.append(M); Bs.append(B)
        n = 64 * len(hs)
        hd = P.add_tensor(hp.H * 2, hp.W * 2, n, _r(n, 32), 0, a3s[0], "heads.dw2")
        wp = np.zeros((hd.pitch, 9), np.int8); wp[:n] = np.concatenate(wqs)
        Mp, Bp = np.zeros(hd.pitch, F), np.zeros(hd.pitch, F)
        Mp[:n], Bp[:n] = np.concatenate(Ms), np.concatenate(Bs)
        P.ops.append(Op("dw", "heads.dw2", dict(in_t=hp.id, out_t=hd.id, in_shift=1, stride=1, wq=wp, C=hd.pitch,
                                                 zx=_zx(a1s[0]), M=Mp, B=Bp, lo=lo)))
        P.taps["heads.act3"] = hd.id
        kmap = {h["out"].name: a3[0] for h, a3 in zip(hs, a3s)}
        hd.act = a3s[0]
        self.emit_pw([(h["out"], 64 * k + np.arange(64)) for k, h in enumerate(hs)], hd, 0, hd.pitch, [None] * len(hs),
                     False, "heads.out", f32=True, kmap=kmap)
        P.cat = hs[0]["classes"]
        P.out_H, P.out_W = hd.H, hd.W
        return P


def build_plan(cfg: NetConfig, state: Dict[str, np.ndarray], in_H: int, in_W: int, offset_mode="round") -> Plan:
    return PlanBuilder(cfg, state, in_H, in_W, offset_mode).build()
