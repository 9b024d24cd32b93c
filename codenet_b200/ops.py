"""Runs single plan ops through the stand-alone C-ABI entry points of libcodenet_b200 on torch CUDA buffers.

This is what the module-level classes of codenet_b200.compat (QuantBnConv2d, QuantBaseNode, ... -- boundary B2 of
SURVEY.md 8(b)) execute: every op of a Plan (codenet_b200.plan) maps to one exported kernel,

    stem    -> cdn_stem_f32_i8        dw      -> cdn_dw3x3_i8
    deform  -> cdn_deform_dw_w4a8     pw      -> cdn_pw_gemm_i8   (int8 NHWC out, or fp32 NCHW planes for head convs)

PyTorch only owns the device memory and the stream.  There is no CPU path: the library refuses to load without an sm_100
device (cdn_check_device).
"""
import ctypes as C

import numpy as np

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def run_op(plan, op, bufs, images=None, want_sval=False):
    """Executes `op`; `bufs` maps tensor ids of `plan` to int8 CUDA tensors [B,H,W,pitch] and receives the output.
    Returns the output tensor (int8 NHWC, or fp32 [B,n,H,W] for an fp32-output 1x1 conv); with want_sval the deformable
    layer also returns its offset scalar map [B,H,W] (fp32)."""
    import torch
    L = _lib.load()
    keep = _lib.Keep()
    a = op.a
    dev = images.device if images is not None else bufs[a["in_t"]].device
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if op.kind == "stem":
            t = plan.tensors[a["out_t"]]
            B = images.shape[0]
            out = torch.empty((B, t.H, t.W, t.pitch), dtype=torch.int8, device=dev)
            rq = keep.requant(a["M"], a["B"], a["lo"])
            _lib.check(L.cdn_stem_f32_i8(_ptr(images), B, a["H"], a["W"], a["stride"], a["pool"], keep.i8(a["wq"]), a["C"],
                                         C.byref(rq), _ptr(out), t.pitch, stream))
            bufs[t.id] = out
            return out
        tin = plan.tensors[a["in_t"]]
        x = bufs[tin.id]
        B = x.shape[0]
        if op.kind in ("dw", "deform"):
            tout = plan.tensors[a["out_t"]]
            out = torch.zeros((B, tout.H, tout.W, tout.pitch), dtype=torch.int8, device=dev)
            rq = keep.requant(a["M"], a["B"], a["lo"])
            H, W = tin.H << a["in_shift"], tin.W << a["in_shift"]
            if op.kind == "dw":
                _lib.check(L.cdn_dw3x3_i8(_ptr(x), tin.pitch, B, H, W, a["in_shift"], a["stride"], keep.i8(a["wq"]), a["C"],
                                          a["zx"], C.byref(rq), _ptr(out), tout.pitch, stream))
                bufs[tout.id] = out
                return out
            sc = keep.deform_scale(a)
            sv = torch.zeros((B, H, W), dtype=torch.float32, device=dev) if want_sval else None
            _lib.check(L.cdn_deform_dw_w4a8(_ptr(x), tin.pitch, B, H, W, a["in_shift"], C.byref(sc), keep.i8(a["wq"]), a["C"],
                                            a["zx"], C.byref(rq), _ptr(out), tout.pitch, _ptr(sv), stream))
            bufs[tout.id] = out
            return (out, sv) if want_sval else out
        if op.kind == "pw":
            pixels = B * tin.H * tin.W
            d = keep.pw_desc(a)
            pas = bufs[a["pass_t"]] if a["pass_t"] >= 0 else None
            pp = plan.tensors[a["pass_t"]].pitch if a["pass_t"] >= 0 else 0
            if a["n_f32"]:
                out = torch.zeros((B, a["n_f32"], tin.H, tin.W), dtype=torch.float32, device=dev)
                _lib.check(L.cdn_pw_gemm_i8(_ptr(x), tin.pitch, pixels, C.byref(d), None, 0, None, 0, _ptr(out), tin.H * tin.W,
                                            stream))
                return out
            tout = plan.tensors[a["out_t"]]
            out = torch.zeros((B, tout.H, tout.W, tout.pitch), dtype=torch.int8, device=dev)
            _lib.check(L.cdn_pw_gemm_i8(_ptr(x), tin.pitch, pixels, C.byref(d), _ptr(pas), pp, _ptr(out), tout.pitch, None, 0,
                                        stream))
            bufs[tout.id] = out
            return out
    raise ValueError(op.kind)


def run_unit(plan, op_pw1, op_dw, op_pw3, bufs):
    """The three ops of a ShuffleNetV2 unit's branch (1x1 conv, depthwise 3x3, 1x1 conv + cat + channel_shuffle) as ONE kernel
    through cdn_shuffle_unit_i8 (QuantBaseNode.forward, quant_modules.py:878-907).  Returns the unit's output tensor, or None when
    the fused kernels do not take this unit's shapes (the caller then runs the three ops one by one)."""
    import torch
    L = _lib.load()
    keep = _lib.Keep()
    a1, ad, a3 = op_pw1.a, op_dw.a, op_pw3.a
    if ad["in_shift"] or a3["pass_t"] < 0 or a1["n_f32"] or a3["n_f32"]:
        return None
    tin, tpass, tout = plan.tensors[a1["in_t"]], plan.tensors[a3["pass_t"]], plan.tensors[a3["out_t"]]
    x, pas = bufs[tin.id], bufs[tpass.id]
    B = x.shape[0]
    with torch.cuda.device(x.device):
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        out = torch.zeros((B, tout.H, tout.W, tout.pitch), dtype=torch.int8, device=x.device)
        rq = keep.requant(ad["M"], ad["B"], ad["lo"])
        d1, d3 = keep.pw_desc(a1), keep.pw_desc(a3)
        r = L.cdn_shuffle_unit_i8(_ptr(x), tin.pitch, B, tin.H, tin.W, ad["stride"], C.byref(d1), keep.i8(ad["wq"]), ad["C"], ad["zx"],
                                  C.byref(rq), C.byref(d3), _ptr(pas), tpass.pitch, _ptr(out), tout.pitch, stream)
    if r != 0:
        if L.cdn_last_error().decode().startswith("unit not fusable"):
            return None
        _lib.check(r)
    bufs[tout.id] = out
    return out


def quantize(x, C_, H, W, scale, zero, pitch):
    """fp32 NCHW CUDA tensor -> int8 NHWC [B,H,W,pitch] on the grid (scale, zero): cdn_quantize_f32_i8."""
    import torch
    L = _lib.load()
    x = x.contiguous().float()
    out = torch.empty((x.shape[0], H, W, pitch), dtype=torch.int8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.cdn_quantize_f32_i8(_ptr(x), x.shape[0], C_, H, W, float(scale), float(zero), _ptr(out), pitch,
                                         C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
    return out


def maxpool3s2(q):
    """nn.MaxPool2d(3, 2, 1) of an int8 NHWC grid: cdn_maxpool3s2_i8."""
    import torch
    L = _lib.load()
    B, H, W, pitch = q.shape
    out = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, pitch), dtype=torch.int8, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(L.cdn_maxpool3s2_i8(_ptr(q), B, H, W, pitch, _ptr(out), C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)))
    return out


def deform_conv_f32(x, offset, weight, stride, padding, dilation, groups, deformable_groups):
    """The general fp32 deformable convolution (the reference's op-level plug-in point): cdn_deform_conv_forward_f32."""
    import torch
    L = _lib.load()
    pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)
    (sH, sW), (pH, pW), (dH, dW) = pair(stride), pair(padding), pair(dilation)
    x, offset, weight = x.contiguous().float(), offset.contiguous().float(), weight.contiguous().float()
    B, Cc, H, W = x.shape
    Co, _, kH, kW = weight.shape
    Ho = (H + 2 * pH - (dH * (kH - 1) + 1)) // sH + 1
    Wo = (W + 2 * pW - (dW * (kW - 1) + 1)) // sW + 1
    out = torch.empty((B, Co, Ho, Wo), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.cdn_deform_conv_forward_f32(_ptr(x), _ptr(weight), _ptr(offset), _ptr(out), B, Cc, H, W, Co, kW, kH, sW, sH,
                                                 pW, pH, dW, dH, groups, deformable_groups, 64,
                                                 C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
    return out
