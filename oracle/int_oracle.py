"""TEST INFRASTRUCTURE ONLY -- the product path never imports this file.

Integer-exact CPU restatement (numpy) of the reference's frozen-range W4A8 CoDeNet forward + ctdet decode.
It is what the CUDA engine must equal bit for bit on the int8 grids and on the top-K indices, and it is pinned
against the vectors the unmodified reference produced in fp64 (tests/golden/, oracle/make_golden.py).

Every formula cites the reference line it restates (paths relative to /root/reference):
  activation quant  s = (2^k-1)/max(M-m,1e-10); z = rint(s*m)+2^(k-1); q = rint(s*x - z); x^ = (q+z)/s
                    portable_quantizer/quantization_utils/quant_utils.py:58-73, :31-50, :191-198 (no clamp);
                    rint = torch.round = half-to-even
  weight quant      per output channel: sigma = (2^(k-1)-1)/max(|min|,|max|,1e-10);
                    wq = clamp(rint(sigma*W), -2^(k-1), 2^(k-1)-1)        quant_utils.py:76-82, :205-223
  BN fold           W' = W*gamma/sqrt(var+eps); b' = (0-mean)*gamma/sqrt(var+eps)+beta
                    portable_quantizer/quant_modules.py:364-372
  graph             quantize_model.py:26-82, quant_modules.py:668-671 (deform block), :878-907 (unit),
                    :1061-1071 (head), lib/models/networks/shufflenetv2_dcn.py:29-34 (shuffle), :314-330
  deformable tap    lib/models/external/src/dcn_deform_conv_cuda_kernel.cu:198-241 and bilinear :83-114;
                    offsets o = anchor*(s-1), lib/models/external/modules/dcn_deform_conv.py:319-330
  decode            lib/models/decode.py:10-16 (_nms), :110-126 (_topk), :474-505 (ctdet_decode)

Arithmetic spec shared with the engine (DESIGN.md "requantisation"): with a = q + z_x the exact integer value
of an activation (real value a/s_x) and acc = sum wq*a the exact integer accumulator,
    q_out = clamp(max(rint(fl(fl(acc*M_c) + B_c)), relu ? -z_out : -inf), -128, 127)
    M_c = s_out/(sigma_c*s_x),  B_c = s_out*b'_c - z_out          (fp64, one rounding per operation)
which differs from the reference's fp64 evaluation order only in the last ulp (a tie within ~1e-13 of .5).
The engine saturates to int8; the reference does not -- `self.saturated` counts the events.
"""
import numpy as np

from codenet_b200.arch import NetConfig, build_graph, act_keys

F = np.float64


def act_params(lo, hi, bits=8):
    lo, hi = F(lo), F(hi)
    # `n / tensor` in torch is tensor.reciprocal() * n (Tensor.__rtruediv__): two roundings, restated as such
    s = (F(1) / max(hi - lo, F(1e-10))) * F(2 ** bits - 1)
    z = np.rint(s * lo) + F(2 ** (bits - 1))
    return s, z


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


def requant(acc, M, B, z_out, relu):
    """acc: int64 [B,C,H,W]; M,B: [C] fp64."""
    t = acc.astype(F) * M.reshape(1, -1, 1, 1)
    t = t + B.reshape(1, -1, 1, 1)
    q = np.rint(t)
    if relu:
        q = np.maximum(q, -z_out)
    return q


def conv_pw(a, wq):
    """a: int64 [B,Cin,H,W] exact activation integers; wq: int64 [Cout,Cin]. Exact (via fp64 BLAS, < 2^53)."""
    B, C, H, W = a.shape
    y = wq.astype(F) @ a.reshape(B, C, H * W).astype(F)
    return np.rint(y).astype(np.int64).reshape(B, -1, H, W)


def conv_dw(a, wq, stride):
    """Depthwise 3x3, pad 1 with REAL zeros (a = 0 outside).  a: int64 [B,C,H,W]; wq: int64 [C,3,3]."""
    B, C, H, W = a.shape
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    p = np.zeros((B, C, H + 2, W + 2), np.int64)
    p[:, :, 1:-1, 1:-1] = a
    acc = np.zeros((B, C, Ho, Wo), np.int64)
    for i in range(3):
        for j in range(3):
            acc += wq[:, i, j].reshape(1, C, 1, 1) * p[:, :, i:i + stride * Ho:stride, j:j + stride * Wo:stride]
    return acc


def conv_dense3(x, wq, stride):
    """Stem: dense 3x3 pad 1 on the fp32 image (values taken exactly as fp64). wq int64 [Cout,3,3,3]."""
    B, C, H, W = x.shape
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    p = np.zeros((B, C, H + 2, W + 2), F)
    p[:, :, 1:-1, 1:-1] = x
    acc = np.zeros((B, wq.shape[0], Ho, Wo), F)
    for c in range(C):
        for i in range(3):
            for j in range(3):
                acc += wq[:, c, i, j].astype(F).reshape(1, -1, 1, 1) * \
                    p[:, c:c + 1, i:i + stride * Ho:stride, j:j + stride * Wo:stride]
    return acc


def maxpool3s2(q):
    B, C, H, W = q.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    p = np.full((B, C, H + 2, W + 2), -np.inf)
    p[:, :, 1:-1, 1:-1] = q
    out = np.full((B, C, Ho, Wo), -np.inf)
    for i in range(3):
        for j in range(3):
            out = np.maximum(out, p[:, :, i:i + 2 * Ho:2, j:j + 2 * Wo:2])
    return out


def deform_dw_int(a, wq, s_int, stride=1):
    """Integer-offset mode: tap (i,j) of output (h,w) reads a[h*st+(i-1)*s, w*st+(j-1)*s] (0 outside)."""
    B, C, H, W = a.shape
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    hh = (np.arange(Ho) * stride).reshape(1, Ho, 1)
    ww = (np.arange(Wo) * stride).reshape(1, 1, Wo)
    s = s_int.astype(np.int64).reshape(B, Ho, Wo)
    acc = np.zeros((B, C, Ho, Wo), np.int64)
    bi = np.arange(B).reshape(B, 1, 1)
    for i in range(3):
        for j in range(3):
            hi, wi = hh + (i - 1) * s, ww + (j - 1) * s
            ok = (hi >= 0) & (hi < H) & (wi >= 0) & (wi < W)
            v = a[bi, :, np.clip(hi, 0, H - 1), np.clip(wi, 0, W - 1)]      # [B,Ho,Wo,C]
            v = np.where(ok[..., None], v, 0).transpose(0, 3, 1, 2)
            acc += wq[:, i, j].reshape(1, C, 1, 1) * v
    return acc


def deform_dw_bilinear(a, wq, s, stride=1):
    """Fractional s: kernel.cu:210-227 range test + :83-114 bilinear, on exact integers a, fp64 result."""
    B, C, H, W = a.shape
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    s = s.reshape(B, Ho, Wo).astype(F)
    d = s - F(1)
    af = a.astype(F)
    bi = np.arange(B).reshape(B, 1, 1)
    acc = np.zeros((B, C, Ho, Wo), F)

    def corner(hi, wi, ok):
        ok = ok & (hi >= 0) & (hi <= H - 1) & (wi >= 0) & (wi <= W - 1)
        v = af[bi, :, np.clip(hi, 0, H - 1), np.clip(wi, 0, W - 1)]
        return np.where(ok[..., None], v, 0.0)

    for i in range(3):
        for j in range(3):
            h_im = (np.arange(Ho) * stride - 1 + i).reshape(1, Ho, 1).astype(F) + F(i - 1) * d
            w_im = (np.arange(Wo) * stride - 1 + j).reshape(1, 1, Wo).astype(F) + F(j - 1) * d
            inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
            hl, wl = np.floor(h_im), np.floor(w_im)
            lh, lw = h_im - hl, w_im - wl
            hh_, hw_ = 1 - lh, 1 - lw
            hl, wl = hl.astype(np.int64), wl.astype(np.int64)
            w1, w2, w3, w4 = hh_ * hw_, hh_ * lw, lh * hw_, lh * lw
            val = w1[..., None] * corner(hl, wl, inside)
            val = val + w2[..., None] * corner(hl, wl + 1, inside)
            val = val + w3[..., None] * corner(hl + 1, wl, inside)
            val = val + w4[..., None] * corner(hl + 1, wl + 1, inside)
            acc += wq[:, i, j].astype(F).reshape(1, C, 1, 1) * val.transpose(0, 3, 1, 2)
    return acc


def shuffle2(x1, x2):
    """cat + channel_shuffle(.,2): out[2k] = x1[k], out[2k+1] = x2[k]  (shufflenetv2_dcn.py:29-34)."""
    B, C, H, W = x1.shape
    return np.stack([x1, x2], axis=2).reshape(B, 2 * C, H, W)


class IntOracle:
    """state: dict in the QUANTISED key space of the reference (numpy arrays), incl. QuantAct x_min/x_max."""

    def __init__(self, cfg: NetConfig, state, offset_mode="round"):
        assert offset_mode in ("round", "bilinear")
        self.cfg, self.g, self.st, self.mode = cfg, build_graph(cfg), state, offset_mode
        self.ak = act_keys(self.g)
        self.saturated = 0
        self.cap = {}

    # -- parameters ------------------------------------------------------------------------------------
    def act(self, label):
        p = self.ak[label]
        return act_params(self.st[p + ".x_min"].reshape(-1)[0], self.st[p + ".x_max"].reshape(-1)[0], self.cfg.a_bit)

    def weights(self, c):
        w = self.st[c.q_conv + ".weight"].astype(F)
        if c.q_bn:
            w, b = fold_bn(w, *(self.st[c.q_bn + "." + f].astype(F) for f in
                                ("weight", "bias", "running_mean", "running_var")))
        elif c.has_bias:
            b = self.st[c.q_conv + ".bias"].astype(F)
        else:
            b = np.zeros(c.cout, F)
        wq, sigma = quant_weight(w, c.w_bit, getattr(self.cfg, "wt_percentile", False))
        return wq, sigma, b

    def sat(self, q):
        self.saturated += int(((q < -128) | (q > 127)).sum())
        return np.clip(q, -128, 127).astype(np.int64)

    def layer(self, c, q_in, in_act, out_act, relu, stride=None):
        """conv + folded BN [+ReLU] + QuantAct on int grids. in_act/out_act: (s,z)."""
        s_x, z_x = in_act
        s_o, z_o = out_act
        wq, sigma, b = self.weights(c)
        a = q_in + np.int64(z_x)
        if c.kind == "pw":
            acc = conv_pw(a, wq.reshape(c.cout, c.cin))
        else:
            acc = conv_dw(a, wq.reshape(c.cout, 3, 3), c.stride if stride is None else stride)
        M = s_o / (sigma * s_x)
        Bc = s_o * b - z_o
        return self.sat(requant(acc, M, Bc, z_o, relu))

    # -- forward ---------------------------------------------------------------------------------------
    def forward(self, x):
        """x: fp32 [B,3,H,W] -> dict(hm_logit, wh, reg fp64 NCHW)."""
        g, cap = self.g, self.cap
        # stem: fp32 image, 8-bit weights (quantize_model.py:26-35)
        wq, sigma, b = self.weights(g.stem)
        s0, z0 = self.act("stem")
        acc = conv_dense3(x.astype(F), wq, g.stem.stride)
        q = self.sat(requant_f(acc, s0 / sigma, s0 * b - z0, z0, True))
        if self.cfg.maxpool:
            q = maxpool3s2(q).astype(np.int64)
        cap["stem"] = q
        cur = (s0, z0)
        for u in g.units:
            r = "layer%d.%d." % (u["stage"], u["unit"])
            cv = u["convs"]
            shared = self.act("layer%d.shared" % u["stage"])
            if u["stride"] == 1:
                half = q.shape[1] // 2
                x1, x2 = q[:, :half], q[:, half:]
            else:
                a4 = self.act(r + "act4")
                x1 = self.layer(cv["dw4"], q, cur, a4, False)
                cap[r + "act4"] = x1
                x1 = self.layer(cv["pw5"], x1, a4, shared, True)
                x2 = q
            a1, a2 = self.act(r + "act1"), self.act(r + "act2")
            x2 = self.layer(cv["pw1"], x2, cur, a1, True)
            cap[r + "act1"] = x2
            x2 = self.layer(cv["dw2"], x2, a1, a2, False)
            cap[r + "act2"] = x2
            x2 = self.layer(cv["pw3"], x2, a2, shared, True)
            q = shuffle2(x1, x2)
            cur = shared
            cap["layer%d.out" % u["stage"]] = q
        a4 = self.act("layer4")
        q = self.layer(g.layer4, q, cur, a4, True)
        cap["layer4"] = q
        cur = a4
        for up in g.ups:
            q, cur = self.up_block(up, q, cur)
            q = np.repeat(np.repeat(q, 2, axis=2), 2, axis=3)          # nearest x2, shufflenetv2_dcn.py:303
        out = {}
        for h in g.heads:
            n = h["name"]
            a1, a3 = self.act(n + ".act1"), self.act(n + ".act3")
            t = self.layer(h["pw1"], q, cur, a1, True)
            cap[n + ".act1"] = t
            t = self.layer(h["dw2"], t, a1, a3, True)
            cap[n + ".act3"] = t
            wq, sigma, b = self.weights(h["out"])
            acc = conv_pw(t + np.int64(a3[1]), wq.reshape(h["classes"], 64))
            y = acc.astype(F) * (F(1) / (sigma * a3[0])).reshape(1, -1, 1, 1)
            out[n] = y + b.reshape(1, -1, 1, 1)
        return out

    def up_block(self, up, q, cur):
        """QuantDeformConvWithOffsetScaleBoundPositive.forward + ReLU + QuantAct (quant_modules.py:668-671,
        quantize_model.py:70-82)."""
        i = up["idx"]
        s_x, z_x = cur
        a = q + np.int64(z_x)
        # offset scalar: Quant_Conv2d C->1 (one scale), Hardtanh, QuantAct
        wq, sigma, b = self.weights(up["scale"])
        acc = conv_pw(a, wq.reshape(1, -1))
        u = acc.astype(F) * (F(1) / (sigma * s_x)).reshape(1, 1, 1, 1) + b.reshape(1, 1, 1, 1)
        u = np.clip(u, F(-self.cfg.offset_bound + 1), F(self.cfg.offset_bound))
        s_s, z_s = self.act("up%d.s" % i)
        qs = np.rint(s_s * u - z_s)
        self.cap["up%d.s" % i] = qs.astype(np.int64)
        s = (qs + z_s) / s_s
        if self.mode == "round":
            s = np.rint(s)
        self.cap["up%d.sval" % i] = s
        # depthwise deformable conv
        wq, sigma, b = self.weights(up["deform"])
        s_d, z_d = self.act("up%d.deform" % i)
        M = s_d / (sigma * s_x)
        Bc = s_d * b - z_d
        if self.mode == "round":
            acc = deform_dw_int(a, wq.reshape(-1, 3, 3), s[:, 0])
            qd = self.sat(requant(acc, M, Bc, z_d, False))
        else:
            accf = deform_dw_bilinear(a, wq.reshape(-1, 3, 3), s[:, 0])
            qd = self.sat(requant_f(accf, M, Bc, z_d, False))
        self.cap["up%d.deform" % i] = qd
        out_act = self.act("up%d.out" % i)
        qo = self.layer(up["channel"], qd, (s_d, z_d), out_act, True)
        self.cap["up%d.out" % i] = qo
        return qo, out_act


def requant_f(accf, M, B, z_out, relu):
    t = accf * M.reshape(1, -1, 1, 1)
    t = t + B.reshape(1, -1, 1, 1)
    q = np.rint(t)
    if relu:
        q = np.maximum(q, -z_out)
    return q


# ---- decode -------------------------------------------------------------------------------------------
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def ctdet_decode(hm_logit, wh, reg=None, K=100):
    """Deterministic restatement of ctdet_decode (decode.py:474-505) on LOGITS (sigmoid is monotone, so
    peaks and order are those of the reference; scores are sigmoid(logit)).

    Peak: value equals the max of its 3x3 neighbourhood, -inf outside (decode.py:10-16).  Order: score
    descending, then class ascending, then spatial index ascending (torch.topk leaves ties unspecified).
    Returns dets [B,K,6] fp64 = x1,y1,x2,y2,score,class and inds [B,K] (class*H*W + spatial index).
    """
    Bn, C, H, W = hm_logit.shape
    p = np.full((Bn, C, H + 2, W + 2), -np.inf)
    p[:, :, 1:-1, 1:-1] = hm_logit
    mx = np.full(hm_logit.shape, -np.inf)
    for i in range(3):
        for j in range(3):
            mx = np.maximum(mx, p[:, :, i:i + H, j:j + W])
    keep = mx == hm_logit
    dets = np.zeros((Bn, K, 6))
    inds = np.zeros((Bn, K), np.int64)
    for b in range(Bn):
        flat = np.where(keep[b], hm_logit[b], -np.inf).reshape(-1)
        order = np.lexsort((np.arange(flat.size), -flat))[:K]            # stable: score desc, then flat index
        assert np.isfinite(flat[order]).all(), "fewer than K peaks"
        cls, sp = order // (H * W), order % (H * W)
        ys, xs = (sp // W).astype(F), (sp % W).astype(F)
        if reg is not None:
            xs = xs + reg[b, 0].reshape(-1)[sp]
            ys = ys + reg[b, 1].reshape(-1)[sp]
        else:
            xs, ys = xs + 0.5, ys + 0.5
        w_, h_ = wh[b, 0].reshape(-1)[sp], wh[b, 1].reshape(-1)[sp]
        dets[b] = np.stack([xs - w_ / 2, ys - h_ / 2, xs + w_ / 2, ys + h_ / 2, sigmoid(flat[order]), cls], 1)
        inds[b] = order
    return dets, inds
