"""TEST INFRASTRUCTURE (never imported by the product): numpy restatement of the operand split of the tensor-core float GEMM
(codenet_b200/csrc/pw_tf32.cu) -- the arithmetic that replaces the fp32 products of the reference's nn.Conv2d(k=1) layers
(lib/models/networks/shufflenetv2_dcn.py:63-99, :209-216, :244-271, :286-300) in EngineF32(gemm="tf32x3").

x = hi + lo with hi = x rounded to TF32 (10 explicit mantissa bits, nearest, ties away from zero = PTX cvt.rna.tf32.f32 for finite
inputs) and lo = the fp32 remainder x - hi (exact) rounded to TF32 the same way.  pack_weights restates cdn_pw_tf32x3_pack: N tiles
of <= 128 output channels, tile columns padded to a multiple of 16, input channels padded to a multiple of 16, every 128-byte row
= 16 channels hi followed by the same 16 channels lo.
"""
import numpy as np


def tf32_rna(x):
    """float32 array -> float32 array with the low 13 mantissa bits cleared after adding half an ulp of TF32 to the magnitude."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    hi = tf32_rna(x)
    lo = tf32_rna((x - hi).astype(np.float32))
    return hi, lo


def tiling(Co):
    nt = (Co + 127) // 128
    per = (Co + nt - 1) // nt
    bn = (per + 15) // 16 * 16
    return nt, per, bn


def pack_weights(w):
    """w [Co][C] float32 -> packed [NT*BN][kpad/16][32] float32 (flattened), cdn_pw_tf32x3_pack's layout."""
    Co, C = w.shape
    nt, per, bn = tiling(Co)
    kpad = (C + 15) // 16 * 16
    full = np.zeros((nt * bn, kpad), np.float32)
    for t in range(nt):
        n = min(per, Co - t * per)
        full[t * bn:t * bn + n, :C] = w[t * per:t * per + n]
    hi, lo = split(full)
    out = np.empty((nt * bn, kpad // 16, 32), np.float32)
    out[:, :, :16] = hi.reshape(nt * bn, kpad // 16, 16)
    out[:, :, 16:] = lo.reshape(nt * bn, kpad // 16, 16)
    return out.reshape(-1)
