"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference's general deformable convolution forward
(lib/models/external/src/dcn_deform_conv_cuda_kernel.cu:189-242 deformable_im2col + the per-group GEMM of
dcn_deform_conv_cuda.cpp:220-235; bilinear sampling :83-114; output size dcn_deform_conv_cuda.cpp:185-190),
and of DeformConvWithOffsetScaleBoundPositive.forward (lib/models/external/modules/dcn_deform_conv.py:323-330).
Pinned against tests/golden/deform_kat.npz."""
import numpy as np


def bilinear(img, h, w):
    """img [C,H,W]; h,w [..] fractional positions already known to satisfy the range test. Returns [C,...]."""
    C, H, W = img.shape
    hl, wl = np.floor(h).astype(np.int64), np.floor(w).astype(np.int64)
    hh_, wh_ = hl + 1, wl + 1
    lh, lw = h - hl, w - wl
    hh, hw = 1 - lh, 1 - lw

    def at(y, x, ok):
        v = img[:, np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)]
        return np.where(ok, v, 0.0)

    v1 = at(hl, wl, (hl >= 0) & (wl >= 0))
    v2 = at(hl, wh_, (hl >= 0) & (wh_ <= W - 1))
    v3 = at(hh_, wl, (hh_ <= H - 1) & (wl >= 0))
    v4 = at(hh_, wh_, (hh_ <= H - 1) & (wh_ <= W - 1))
    return hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4


def deform_conv(x, offset, weight, stride=1, pad=1, dil=1, groups=1, dg=1):
    """x [B,C,H,W]; offset [B,2*kH*kW*dg,Ho,Wo] (dy,dx interleaved per tap); weight [Co,C/groups,kH,kW]."""
    B, C, H, W = x.shape
    Co, cpg, kH, kW = weight.shape
    Ho = (H + 2 * pad - (dil * (kH - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kW - 1) + 1)) // stride + 1
    out = np.zeros((B, Co, Ho, Wo), x.dtype)
    cpdg = C // dg
    copg = Co // groups
    hs = (np.arange(Ho) * stride - pad).reshape(Ho, 1)
    ws = (np.arange(Wo) * stride - pad).reshape(1, Wo)
    for b in range(B):
        for g in range(dg):
            img = x[b, g * cpdg:(g + 1) * cpdg]
            cols = np.zeros((cpdg, kH * kW, Ho, Wo), x.dtype)
            for i in range(kH):
                for j in range(kW):
                    t = i * kW + j
                    h_im = hs + i * dil + offset[b, (g * kH * kW + t) * 2]
                    w_im = ws + j * dil + offset[b, (g * kH * kW + t) * 2 + 1]
                    ok = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
                    cols[:, t] = np.where(ok, bilinear(img, h_im, w_im), 0.0)
            for co in range(Co):
                grp = co // copg
                for cl in range(cpg):
                    c = grp * cpg + cl
                    if c // cpdg != g:
                        continue
                    out[b, co] += np.tensordot(weight[co, cl].reshape(-1), cols[c - g * cpdg], axes=(0, 0))
    return out


ANCHOR = np.array([-1, -1, -1, 0, -1, 1, 0, -1, 0, 0, 0, 1, 1, -1, 1, 0, 1, 1], np.float64).reshape(1, 18, 1, 1)


def codesigned_module(x, w_scale, b_scale, w_dw, stride, bound, w_channel=None):
    """s = Hardtanh(conv1x1_{C->1, stride}(x)); o = anchor*(s-1); y = deform_conv(x, o, w_dw, groups=C)."""
    xs = x[:, :, ::stride, ::stride]
    s = np.tensordot(xs, w_scale.reshape(-1), axes=(1, 0))[:, None] + b_scale.reshape(1, 1, 1, 1)
    s = np.clip(s, -bound + 1, bound)
    o = ANCHOR * (s - 1)
    y = deform_conv(x, o, w_dw, stride, 1, 1, groups=x.shape[1], dg=1)
    if w_channel is not None:
        y = np.tensordot(w_channel.reshape(w_channel.shape[0], -1), y, axes=(1, 1)).transpose(1, 0, 2, 3)
    return y
