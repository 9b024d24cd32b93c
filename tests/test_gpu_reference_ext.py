"""The reference's OWN CUDA extension (lib/models/external/src/dcn_deform_conv_cuda{.cpp,_kernel.cu}, unmodified, built
for sm_100a by oracle/build_ref.py) executed on the B200 as the checker:

  * pins the numpy oracle (oracle/deform_ref.py) against the real reference kernel, not only against the torchvision
    stand-in the CPU fixtures were generated with;
  * checks `cdn_deform_conv_forward_f32` -- the drop-in for `deform_conv_forward_cuda`
    (dcn_deform_conv_cuda.cpp:151-258, called as in functions/dcn_deform_conv.py:51-56, W-before-H argument order);
  * checks the fused W4A8 integer-offset layer against the reference op fed the offsets anchor*(s-1) the co-designed module
    builds (modules/dcn_deform_conv.py:319-330): with integer s the reference's bilinear gather is exact.
"""
import ctypes as C

import numpy as np
import pytest

from codenet_b200 import _lib
from oracle import build_ref, deform_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_ext():
    if not build_ref.available():
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py in the build container)")
    return build_ref.load()


def ref_forward(ext, x, off, w, stride, pad, dil, groups, dg):
    """The call of DeformConvFunction.forward (functions/dcn_deform_conv.py:36-56)."""
    import torch
    B, Cc, H, W = x.shape
    Co, _, kH, kW = w.shape
    Ho = (H + 2 * pad - (dil * (kH - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kW - 1) + 1)) // stride + 1
    out = x.new_empty((B, Co, Ho, Wo))
    step = min(64, B)
    assert B % step == 0
    r = ext.deform_conv_forward_cuda(x, w, off, out, x.new_empty(0), x.new_empty(0), kW, kH, stride, stride, pad, pad, dil, dil,
                                     groups, dg, step)
    assert r == 1
    torch.cuda.synchronize()
    return out


def ours_forward(x, off, w, stride, pad, dil, groups, dg):
    import torch
    from gpu_util import ptr, stream
    L = _lib.load()
    B, Cc, H, W = x.shape
    Co, _, kH, kW = w.shape
    Ho = (H + 2 * pad - (dil * (kH - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kW - 1) + 1)) // stride + 1
    out = torch.zeros((B, Co, Ho, Wo), dtype=torch.float32, device="cuda")
    _lib.check(L.cdn_deform_conv_forward_f32(ptr(x), ptr(w), ptr(off), ptr(out), B, Cc, H, W, Co, kW, kH, stride, stride,
                                             pad, pad, dil, dil, groups, dg, 64, stream()))
    torch.cuda.synchronize()
    return out


CASES = [  # B, C, H, W, Co, stride, groups, dg
    (2, 16, 12, 14, 16, 1, 16, 1),          # depthwise (the CoDeNet use)
    (4, 32, 16, 16, 32, 2, 32, 1),          # depthwise stride 2
    (2, 8, 10, 9, 12, 1, 1, 1),             # dense
    (2, 8, 9, 11, 8, 1, 2, 2),              # grouped, two deformable groups
    (64, 24, 8, 8, 24, 1, 24, 1),           # a full im2col_step chunk
]


@pytest.mark.parametrize("B,Cc,H,W,Co,stride,groups,dg", CASES)
def test_general_op_matches_reference_extension(ref_ext, B, Cc, H, W, Co, stride, groups, dg):
    import torch
    from gpu_util import dev
    rng = np.random.default_rng(B * 100 + Cc + stride)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    x = rng.standard_normal((B, Cc, H, W)).astype(np.float32)
    off = rng.uniform(-3.2, 3.2, (B, 18 * dg, Ho, Wo)).astype(np.float32)
    w = rng.standard_normal((Co, Cc // groups, 3, 3)).astype(np.float32)
    tx, to, tw = dev(x), dev(off), dev(w)
    want = ref_forward(ref_ext, tx, to, tw, stride, 1, 1, groups, dg).cpu().numpy()
    got = ours_forward(tx, to, tw, stride, 1, 1, groups, dg).cpu().numpy()
    scale = max(1.0, float(np.abs(want).max()))
    assert np.abs(got - want).max() <= 1e-4 * scale           # fp32 on both sides, different summation order
    if B <= 4:                                                # and the numpy oracle agrees with the real reference kernel
        ora = deform_ref.deform_conv(x.astype(np.float64), off.astype(np.float64), w.astype(np.float64), stride, 1, 1, groups, dg)
        assert np.abs(ora - want).max() <= 1e-4 * scale


def test_integer_offsets_are_exact_on_both(ref_ext):
    """Integer data, integer weights, integer offsets: every product and sum is exact in fp32, so the reference kernel,
    our kernel and the oracle must agree bit for bit (this is the arithmetic of the W4A8 integer-offset path)."""
    import torch
    from gpu_util import dev
    rng = np.random.default_rng(5)
    B, Cc, H, W = 4, 32, 16, 16
    x = rng.integers(-128, 128, (B, Cc, H, W)).astype(np.float32)
    s = rng.integers(-7, 9, (B, 1, H, W)).astype(np.float32)
    off = (deform_ref.ANCHOR * (s - 1)).astype(np.float32)
    w = rng.integers(-8, 8, (Cc, 1, 3, 3)).astype(np.float32)
    tx, to, tw = dev(x), dev(off), dev(w)
    want = ref_forward(ref_ext, tx, to, tw, 1, 1, 1, Cc, 1).cpu().numpy()
    got = ours_forward(tx, to, tw, 1, 1, 1, Cc, 1).cpu().numpy()
    ora = deform_ref.deform_conv(x.astype(np.float64), off.astype(np.float64), w.astype(np.float64), 1, 1, 1, Cc, 1)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(ora.astype(np.float32), want)


@pytest.mark.parametrize("Cc,H,W,bound,shift", [(128, 16, 16, 8, 0), (256, 16, 16, 4, 1), (24, 20, 12, 2, 0)])
def test_fused_w4a8_layer_matches_reference_extension(ref_ext, Cc, H, W, bound, shift):
    """cdn_deform_dw_w4a8 (integer offsets) == requantisation of the REFERENCE op's output when the reference op is given
    the real-valued inputs (q + zx), the integer weights and the offsets anchor*(s-1) for the s our kernel reports."""
    import torch
    from gpu_util import dev, ptr, stream
    rng = np.random.default_rng(Cc + bound)
    L = _lib.load()
    keep = _lib.Keep()
    B = 2
    pitch = (Cc + 31) // 32 * 32
    Hs, Ws = H >> shift, W >> shift
    zx = int(rng.integers(100, 129))
    q = np.zeros((B, Hs, Ws, pitch), np.int8); q[..., :Cc] = rng.integers(-128, 128, (B, Hs, Ws, Cc))
    wq = np.zeros((pitch, 9), np.int8); wq[:Cc] = rng.integers(-8, 8, (Cc, 9))
    ws = np.zeros(pitch, np.int8); ws[:Cc] = rng.integers(-8, 8, Cc)
    M = np.zeros(pitch); Bc = np.zeros(pitch)
    M[:Cc] = rng.uniform(0.004, 0.02, Cc); Bc[:Cc] = rng.uniform(-20, 20, Cc)
    # scale conv constants chosen so that u spans a bit more than [-bound+1, bound]
    acc_s = (q[..., :Cc].astype(np.int64) * ws[:Cc].astype(np.int64)).sum(-1) + zx * int(ws[:Cc].astype(np.int64).sum())
    span = max(1.0, float(np.abs(acc_s - acc_s.mean()).max()))
    Ms = (bound + 1.0) / span
    ss = 255.0 / (2 * bound - 1)
    a = dict(ws=ws, Ms=Ms, bs=0.5 - Ms * float(acc_s.mean()), ss=ss, zs=float(np.rint(ss * (-bound + 1)) + 128), bound=bound, mode=0)
    sc = keep.deform_scale(a)
    rq = keep.requant(M, Bc, -128)
    tq = dev(q)
    out = torch.zeros((B, H, W, pitch), dtype=torch.int8, device="cuda")
    sv = torch.zeros((B, H, W), dtype=torch.float32, device="cuda")
    _lib.check(L.cdn_deform_dw_w4a8(ptr(tq), pitch, B, H, W, shift, C.byref(sc), keep.i8(wq), pitch, zx, C.byref(rq), ptr(out),
                                    pitch, ptr(sv), stream()))
    torch.cuda.synchronize()
    s = sv.cpu().numpy()[:, None]
    assert len(np.unique(s)) >= min(4, 2 * bound) and s.min() >= -bound + 1 and s.max() <= bound
    # reference op on the real values: nearest x2 upsample of (q + zx) as NCHW fp32
    real = (q[..., :Cc].astype(np.float32) + zx).transpose(0, 3, 1, 2)
    if shift:
        real = real.repeat(2, axis=2).repeat(2, axis=3)
    off = (deform_ref.ANCHOR * (s.astype(np.float64) - 1)).astype(np.float32)
    wf = wq[:Cc].reshape(Cc, 1, 3, 3).astype(np.float32)
    acc = ref_forward(ref_ext, dev(np.ascontiguousarray(real)), dev(off), dev(wf), 1, 1, 1, Cc, 1).cpu().numpy().astype(np.float64)
    assert np.all(acc == np.rint(acc))                         # exact integers out of the reference kernel
    want = np.clip(np.rint(acc * M[:Cc].reshape(1, Cc, 1, 1) + Bc[:Cc].reshape(1, Cc, 1, 1)), -128, 127).astype(np.int8)
    got = out.cpu().numpy()[..., :Cc].transpose(0, 3, 1, 2)
    np.testing.assert_array_equal(got, want)
