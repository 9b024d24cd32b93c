"""GPU box only: the reference's own CUDA path for the deformable layers, timed on the B200 next to ours.

Reference arm = what `DeformConvWithOffsetScaleBoundPositive.forward` launches for the depthwise deformable conv itself
(lib/models/external/modules/dcn_deform_conv.py:323-330 -> functions/dcn_deform_conv.py:51-56 ->
dcn_deform_conv_cuda.cpp:151-258): the unmodified extension built by oracle/build_ref.py, fp32 NCHW, im2col_step 64, given
precomputed offsets (its scale conv / Hardtanh / quantisers / requantisation are NOT in the timed region, ours are).
Ours = the fused W4A8 layer (`cdn_deform_layer_run`, integer offsets) on the same shapes.  Writes gpurun_out/ref_ext_timing.json.
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codenet_b200 import _lib  # noqa: E402
from oracle import build_ref  # noqa: E402

ANCHOR = np.array([-1, -1, -1, 0, -1, 1, 0, -1, 0, 0, 0, 1, 1, -1, 1, 0, 1, 1], np.float32).reshape(1, 18, 1, 1)


def time_it(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ext = build_ref.load()
    L = _lib.load()
    rng = np.random.default_rng(0)
    out = []
    B = 64
    for Cc, H, shift in ((1024, 16, 0), (256, 32, 1), (128, 64, 1)):           # the three layers of CoDeNet1x at 512x512
        bound = 8
        pitch = (Cc + 31) // 32 * 32
        Hs = H >> shift
        # ---- reference extension: fp32 NCHW at the sampling resolution -------------------------------------------------
        x = torch.randn(B, Cc, H, H, device="cuda")
        s = torch.from_numpy(rng.integers(-bound + 1, bound + 1, (B, 1, H, H)).astype(np.float32)).cuda()
        off = (torch.from_numpy(ANCHOR).cuda() * (s - 1)).contiguous()
        w = torch.randn(Cc, 1, 3, 3, device="cuda")
        y = x.new_empty((B, Cc, H, H))
        b0, b1 = x.new_empty(0), x.new_empty(0)
        ms_ref = time_it(lambda: ext.deform_conv_forward_cuda(x, w, off, y, b0, b1, 3, 3, 1, 1, 1, 1, 1, 1, Cc, 1, 64), iters=3)
        # ---- ours: fused W4A8 layer ---------------------------------------------------------------------------------
        keep = _lib.Keep()
        q = torch.from_numpy(rng.integers(-128, 128, (B, Hs, Hs, pitch)).astype(np.int8)).cuda()
        wq = np.zeros((pitch, 9), np.int8); wq[:Cc] = rng.integers(-8, 8, (Cc, 9))
        ws = np.zeros(pitch, np.int8); ws[:Cc] = rng.integers(-8, 8, Cc)
        M = np.zeros(pitch); Bc = np.zeros(pitch)
        M[:Cc] = rng.uniform(0.004, 0.02, Cc); Bc[:Cc] = rng.uniform(-20, 20, Cc)
        ss = 255.0 / (2 * bound - 1)
        a = dict(ws=ws, Ms=float((bound + 1.0) / (np.abs(ws).sum() * 40.0)), bs=0.5, ss=ss, zs=float(np.rint(ss * (-bound + 1)) + 128),
                 bound=bound, mode=0)
        sc = keep.deform_scale(a)
        rq = keep.requant(M, Bc, -128)
        h = C.c_void_p()
        _lib.check(L.cdn_deform_layer_create(C.byref(h), C.byref(sc), keep.i8(wq), pitch, pitch, 128, C.byref(rq)))
        o = torch.zeros((B, H, H, pitch), dtype=torch.int8, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        ms_ours = time_it(lambda: _lib.check(L.cdn_deform_layer_run(h, C.c_void_p(q.data_ptr()), pitch, B, H, H, shift,
                                                                   C.c_void_p(o.data_ptr()), pitch, None, st)), iters=20)
        L.cdn_deform_layer_destroy(h)
        out.append({"C": Cc, "H": H, "batch": B, "in_shift": shift, "reference_ext_fp32_ms": round(ms_ref, 3),
                    "ours_fused_w4a8_ms": round(ms_ours, 4), "speedup": round(ms_ref / ms_ours, 1),
                    "reference_columns_MB": round(36.0 * Cc * H * H * 64 / 1e6, 1)})
        print(out[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"what": __doc__.strip().splitlines()[0], "layers": out}, open("gpurun_out/ref_ext_timing.json", "w"), indent=1)


if __name__ == "__main__":
    main()
