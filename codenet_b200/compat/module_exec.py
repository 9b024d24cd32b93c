"""Module-level execution (boundary B2): turns ONE Quant* module call into a few plan ops and runs them.

Every compat.Quant* module's forward builds a miniature Plan with the same emitters the whole-network compiler uses
(codenet_b200.plan.PlanBuilder: BN folding, per-channel weight quantisation, exact requantisation constants, HALF layout,
chunk table for cat + channel_shuffle) from the module's own parameters, and executes its ops through the stand-alone C-ABI
entry points (codenet_b200.ops).  Results are therefore bit-identical to the compiled engine's, layer by layer.
Like the reference, weights are re-quantised on every call; this path is for users who keep the reference's module structure
(a partially quantised model, a single deformable block); the fast path is the compiled Engine.
"""
from types import SimpleNamespace

import numpy as np
import torch

from .. import ops
from ..arch import ConvSpec
from ..plan import Plan, PlanBuilder, act_params
from .qtensor import QTensor, PendingConv


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def act_of(qact):
    """(scale, zero point) of a compat.QuantAct with a FROZEN range."""
    if qact.full_precision_flag:
        raise NotImplementedError("QuantAct(full_precision_flag=True) leaves the int8 path; not built")
    if qact.running_stat:
        raise RuntimeError("QuantAct range is still a running statistic (the reference updates it on every forward, even in "
                           "eval mode); load calibrated ranges, then freeze_ranges(model) / act.set_range(lo, hi)")
    return act_params(float(qact.x_min.reshape(-1)[0]), float(qact.x_max.reshape(-1)[0]), qact.activation_bit)


def find_act(seq):
    """the QuantAct of an nn.Sequential(ReLU | Hardtanh, QuantAct) member (quantize_model.py) or a bare QuantAct"""
    from .quant_modules import QuantAct
    if isinstance(seq, QuantAct):
        return seq
    for m in seq:
        if isinstance(m, QuantAct):
            return m
    raise ValueError("no QuantAct in %r" % (seq,))


class Mini(PlanBuilder):
    """PlanBuilder over an ad-hoc state dict filled from live modules."""

    def __init__(self, offset_mode="bilinear", wt_percentile=False, offset_bound=8):
        self.cfg = SimpleNamespace(wt_percentile=bool(wt_percentile), a_bit=8, offset_bound=int(offset_bound))
        self.st, self.g, self.ak = {}, None, {}
        self.plan = Plan(self.cfg, 0, 0, offset_mode)
        self.bufs = {}
        self._n = 0

    def conv(self, kind, conv_w, bn=None, bias=None, w_bit=4, stride=1, name=None):
        """registers a conv's parameters and returns its ConvSpec; conv_w: weight tensor [Co, Ci/g, k, k]"""
        self._n += 1
        name = name or "m%d" % self._n
        co, cig, k, _ = conv_w.shape
        self.st[name + ".weight"] = _np(conv_w)
        if bn is not None:
            for f in ("weight", "bias", "running_mean", "running_var"):
                self.st[name + ".bn." + f] = _np(getattr(bn, f))
            if abs(bn.eps - 1e-5) > 1e-12:
                raise NotImplementedError("BatchNorm eps %g: the plan compiler folds with the reference's 1e-5" % bn.eps)
        if bias is not None:
            if bn is not None:
                raise NotImplementedError("a conv with both a bias and a BatchNorm is not on the CoDeNet path")
            self.st[name + ".bias"] = _np(bias)
        groups = co if kind in ("dw", "deform_dw") else 1
        return ConvSpec(name, kind, cig * groups if groups == 1 else co, co, k, stride, groups, "", "", name,
                        name + ".bn" if bn is not None else "", bias is not None, w_bit)

    def bind(self, x: QTensor, name="in"):
        t = x.spec(self.plan, name)
        self.bufs[t.id] = x.q
        return t

    def run(self, images=None):
        out = None
        P = self.plan.ops
        i = 0
        while i < len(P):
            # a unit's branch -- 1x1 conv, depthwise 3x3, interleaving 1x1 conv -- runs as ONE kernel where the fused kernels take it
            if (i + 2 < len(P) and P[i].kind == "pw" and P[i + 1].kind == "dw" and P[i + 2].kind == "pw" and
                    P[i + 1].a["in_t"] == P[i].a["out_t"] and P[i + 2].a["in_t"] == P[i + 1].a["out_t"] and P[i + 2].a["pass_t"] >= 0):
                out = ops.run_unit(self.plan, P[i], P[i + 1], P[i + 2], self.bufs)
                if out is not None:
                    i += 3
                    continue
            out = ops.run_op(self.plan, P[i], self.bufs, images=images)
            i += 1
        return out

    def result(self, t, up=0):
        return QTensor(self.bufs[t.id], t.C, t.act, t.half, up)


def wt_pct(module):
    return bool(getattr(module, "weight_percentile", False))


# ---- QuantBnConv2d / Quant_Conv2d -------------------------------------------------------------------------------------------
def conv_kind(conv, x):
    k = conv.kernel_size[0] if isinstance(conv.kernel_size, tuple) else conv.kernel_size
    if torch.is_tensor(x):
        return "stem"
    if k == 1 and conv.groups == 1:
        return "pw"
    if k == 3 and conv.groups == conv.in_channels == conv.out_channels:
        return "dw"
    raise NotImplementedError("conv %dx%d groups=%d is not on the CoDeNet path (1x1 dense / 3x3 depthwise / 3x3 stem)" % (k, k, conv.groups))


def fused_conv(conv, bn, w_bit, percentile, x, out_act, relu):
    """conv (+ folded BN) [+ ReLU] + QuantAct as one kernel -> QTensor"""
    kind = conv_kind(conv, x)
    stride = conv.stride[0] if isinstance(conv.stride, tuple) else conv.stride
    m = Mini(wt_percentile=percentile)
    if kind == "stem":
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
            raise RuntimeError("the stem conv takes a CUDA fp32 [B,3,H,W] image (codenet_b200 has no CPU path)")
        c = m.conv("stem", conv.weight, bn, conv.bias, w_bit, stride)
        wq, sigma, b = m.weights(c)
        s0, z0 = out_act
        H, W = x.shape[2], x.shape[3]
        t = m.plan.add_tensor((H - 1) // stride + 1, (W - 1) // stride + 1, c.cout, 32, 0, out_act, "stem")
        from ..plan import Op, F
        m.plan.ops.append(Op("stem", "stem", dict(out_t=t.id, H=H, W=W, stride=stride, pool=0, wq=wq.astype(np.int8), C=c.cout,
                                                    M=F(s0) / sigma, B=F(s0) * b - F(z0), lo=max(-128, int(-z0)) if relu else -128)))
        m.run(images=x.contiguous())
        return m.result(t)
    tin = m.bind(x)
    if kind == "pw":
        c = m.conv("pw", conv.weight, bn, conv.bias, w_bit)
        tout = m.emit_pw([(c, x.phys(np.arange(x.C)))], tin, 0, tin.pitch, [out_act], relu, "pw")
        m.run()
        return m.result(tout, x.up)                    # pointwise ops commute with the pending nearest upsample
    c = m.conv("dw", conv.weight, bn, conv.bias, w_bit, stride)
    tout = m.emit_dw(c, tin, out_act, relu, stride, x.up, "dw")
    m.run()
    return m.result(tout)


def float_conv(conv, bn, w_bit, percentile, x, kmap_scale=None):
    """1x1 conv with fp32 output (no QuantAct behind it: head output convs, the offset-scale conv) -> fp32 NCHW"""
    if conv_kind(conv, x) != "pw":
        raise NotImplementedError("an fp32 result is available for 1x1 convs only; close other convs with a QuantAct")
    m = Mini(wt_percentile=percentile)
    tin = m.bind(x)
    c = m.conv("head_out", conv.weight, bn, conv.bias, w_bit)
    m.emit_pw([(c, x.phys(np.arange(x.C)))], tin, 0, tin.pitch, [None], False, "out", f32=True)
    y = m.run()
    for _ in range(x.up):
        y = y.repeat_interleave(2, 2).repeat_interleave(2, 3)
    return y


# ---- compound modules ----------------------------------------------------------------------------------------------------------
def deform_block(mod, x: QTensor):
    """QuantDeformConvWithOffsetScaleBoundPositive.forward up to quant_identity_deform: the fused kernel cdn_deform_dw_w4a8"""
    from ..plan import F
    mode = getattr(mod, "offset_mode", "bilinear")
    m = Mini(offset_mode=mode, wt_percentile=mod.weight_percentile, offset_bound=mod.offset_bound)
    tin = m.bind(x)
    cs = m.conv("scale", mod.quant_conv_scale.weight, None, mod.quant_conv_scale.bias, mod.weight_bit)
    cd = m.conv("deform_dw", mod.quant_deform_conv.weight, None, None, mod.weight_bit)
    wq_s, sigma_s, b_s = m.weights(cs)
    ss, zs = act_of(find_act(mod.quant_act))
    ws = np.zeros(tin.pitch, np.int8)
    ws[x.phys(np.arange(x.C))] = wq_s.reshape(-1)
    extra = dict(ws=ws, Ms=float(F(1) / (sigma_s[0] * F(x.act[0]))), bs=float(b_s[0]), ss=ss, zs=zs, bound=int(mod.offset_bound),
                 mode=0 if mode == "round" else 1)
    tout = m.emit_dw(cd, tin, act_of(mod.quant_identity_deform), False, 1, x.up, "deform", extra=extra)
    m.run()
    return m.result(tout)


def base_node(mod, x: QTensor):
    """QuantBaseNode.forward (quant_modules.py:878-907): the ShuffleNetV2 unit with split / cat / channel_shuffle folded into
    the last 1x1 conv's epilogue; the output is in the HALF layout the next unit reads."""
    if x.up:
        raise NotImplementedError("QuantBaseNode behind an upsample is not on the path")
    from ..plan import _r
    pct = mod.weight_percentile
    m = Mini(wt_percentile=pct)
    tin = m.bind(x)
    shared = act_of(mod.quant_act)
    a1, a2 = act_of(mod.quant_act1), act_of(mod.quant_act2)
    b = mod.weight_bit
    cb = lambda kind, qm, stride=1: m.conv(kind, qm.conv.weight, qm.bn, qm.conv.bias, b, stride)
    pw1, dw2, pw3 = cb("pw", mod.quant_convbn1), cb("dw", mod.quant_convbn2, mod.stride), cb("pw", mod.quant_convbn3)
    half = pw3.cout
    Hp = _r(half, 32)
    if mod.stride == 2:
        kpos = x.phys(np.arange(x.C))
        a4 = act_of(mod.quant_act4)
        dw4, pw5 = cb("dw", mod.quant_convbn4, 2), cb("pw", mod.quant_convbn5)
        d4 = m.emit_dw(dw4, tin, a4, False, 2, 0, "dw4")
        x1 = m.emit_pw([(pw5, kpos)], d4, 0, tin.pitch, [shared], True, "pw5")
        c1 = m.emit_pw([(pw1, kpos)], tin, 0, tin.pitch, [a1], True, "pw1")
        d2 = m.emit_dw(dw2, c1, a2, False, 2, 0, "dw2")
        out = m.plan.add_tensor(d2.H, d2.W, 2 * half, 2 * Hp, Hp, shared, "out")
        m.emit_pw([(pw3, np.arange(half))], d2, 0, d2.pitch, [shared], True, "pw3", out=out, interleave_with=(x1, 0))
    else:
        if not (x.half == Hp and x.C == 2 * half):
            raise RuntimeError("a stride-1 unit reads the HALF layout its predecessor wrote (got half=%d, C=%d)" % (x.half, x.C))
        if not x.same_grid(shared):
            raise RuntimeError("the pass-through half must already be on the stage's shared grid (quantize_model.py:40)")
        c1 = m.emit_pw([(pw1, np.arange(half))], tin, Hp, Hp, [a1], True, "pw1")
        d2 = m.emit_dw(dw2, c1, a2, False, 1, 0, "dw2")
        out = m.plan.add_tensor(tin.H, tin.W, 2 * half, 2 * Hp, Hp, shared, "out")
        m.emit_pw([(pw3, np.arange(half))], d2, 0, d2.pitch, [shared], True, "pw3", out=out, interleave_with=(tin, 0))
    m.run()
    return m.result(out)


def head_node(mod, x: QTensor):
    """QuantDepthwiseNode.forward (quant_modules.py:1061-1071): pw + ReLU + Q, dw + ReLU + Q (through the pending upsample),
    1x1 output conv with fp32 bias -> fp32 NCHW"""
    m = Mini(wt_percentile=mod.weight_percentile)
    tin = m.bind(x)
    a1, a3 = act_of(find_act(mod.quant_act1)), act_of(find_act(mod.quant_act3))
    b = mod.weight_bit
    pw1 = m.conv("pw", mod.quant_convbn1.conv.weight, mod.quant_convbn1.bn, mod.quant_convbn1.conv.bias, b)
    dw2 = m.conv("dw", mod.quant_convbn2.conv.weight, mod.quant_convbn2.bn, mod.quant_convbn2.conv.bias, b)
    out = m.conv("head_out", mod.quant_conv.weight, None, mod.quant_conv.bias, b)
    hp = m.emit_pw([(pw1, x.phys(np.arange(x.C)))], tin, 0, tin.pitch, [a1], True, "pw1")
    hd = m.emit_dw(dw2, hp, a3, True, 1, x.up, "dw2")
    m.emit_pw([(out, np.arange(hd.C))], hd, 0, hd.pitch, [None], False, "out", f32=True)
    return m.run()
