"""Thin helpers for calling the C ABI on torch CUDA buffers from the GPU tests."""
import ctypes as C

import numpy as np
import torch

from codenet_b200 import _lib


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run_op(plan, op, T, images=None, sval=False):
    """Runs ONE plan op through its stand-alone C-ABI entry point; inputs are taken from the dict of simulated
    tensors T (int64 [B,H,W,pitch]).  Returns the produced tensor as numpy."""
    L = _lib.load()
    keep = _lib.Keep()
    a = op.a
    if op.kind == "stem":
        t = plan.tensors[a["out_t"]]
        B = images.shape[0]
        out = torch.zeros((B, t.H, t.W, t.pitch), dtype=torch.int8, device="cuda")
        rq = keep.requant(a["M"], a["B"], a["lo"])
        img = dev(images.astype(np.float32))
        _lib.check(L.cdn_stem_f32_i8(ptr(img), B, a["H"], a["W"], a["stride"], a["pool"], keep.i8(a["wq"]), a["C"],
                                     C.byref(rq), ptr(out), t.pitch, stream()))
        return out.cpu().numpy()
    if op.kind in ("dw", "deform"):
        tin, tout = plan.tensors[a["in_t"]], plan.tensors[a["out_t"]]
        x = dev(T[tin.id].astype(np.int8))
        B = x.shape[0]
        out = torch.zeros((B, tout.H, tout.W, tout.pitch), dtype=torch.int8, device="cuda")
        rq = keep.requant(a["M"], a["B"], a["lo"])
        H, W = tin.H << a["in_shift"], tin.W << a["in_shift"]
        if op.kind == "dw":
            _lib.check(L.cdn_dw3x3_i8(ptr(x), tin.pitch, B, H, W, a["in_shift"], a["stride"], keep.i8(a["wq"]), a["C"],
                                      a["zx"], C.byref(rq), ptr(out), tout.pitch, stream()))
            return out.cpu().numpy()
        sc = keep.deform_scale(a)
        sv = torch.zeros((B, H, W), dtype=torch.float32, device="cuda") if sval else None
        _lib.check(L.cdn_deform_dw_w4a8(ptr(x), tin.pitch, B, H, W, a["in_shift"], C.byref(sc), keep.i8(a["wq"]), a["C"],
                                        a["zx"], C.byref(rq), ptr(out), tout.pitch, ptr(sv), stream()))
        return (out.cpu().numpy(), sv.cpu().numpy()) if sval else out.cpu().numpy()
    if op.kind == "pw":
        tin = plan.tensors[a["in_t"]]
        x = dev(T[tin.id].astype(np.int8))
        B = x.shape[0]
        pixels = B * tin.H * tin.W
        d = keep.pw_desc(a)
        pas = dev(T[a["pass_t"]].astype(np.int8)) if a["pass_t"] >= 0 else None
        pp = plan.tensors[a["pass_t"]].pitch if a["pass_t"] >= 0 else 0
        if a["n_f32"]:
            out = torch.zeros((B, a["n_f32"], tin.H, tin.W), dtype=torch.float32, device="cuda")
            _lib.check(L.cdn_pw_gemm_i8(ptr(x), tin.pitch, pixels, C.byref(d), None, 0, None, 0, ptr(out), tin.H * tin.W,
                                        stream()))
            return out.cpu().numpy()
        tout = plan.tensors[a["out_t"]]
        out = torch.zeros((B, tout.H, tout.W, tout.pitch), dtype=torch.int8, device="cuda")
        _lib.check(L.cdn_pw_gemm_i8(ptr(x), tin.pitch, pixels, C.byref(d), ptr(pas), pp, ptr(out), tout.pitch, None, 0,
                                    stream()))
        return out.cpu().numpy()
    raise ValueError(op.kind)
