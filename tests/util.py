"""Helpers shared by the parity tests."""
import numpy as np


def _plateau_mask(hm):
    """[C,H,W] bool: the pixel has an 8-neighbour of the same class with EXACTLY the same logit.  The reference evaluates
    `hmax == heat` (decode.py:10-16) on floating-point head outputs: two neighbours whose exact values are equal may differ
    by an ulp there (summation order of its conv), and then only one of them survives as a peak.  Which one is decided by the
    reference's rounding noise, not by the arithmetic of the path -- the exact oracle keeps both."""
    C, H, W = hm.shape
    p = np.full((C, H + 2, W + 2), np.nan)
    p[:, 1:-1, 1:-1] = hm
    m = np.zeros(hm.shape, bool)
    for i in range(3):
        for j in range(3):
            if (i, j) != (1, 1):
                m |= p[:, i:i + H, j:j + W] == hm
    return m


def assert_dets_match_tie_aware(dets, ref, ref_more=None, rtol=1e-5, atol=1e-5, hm=None, more_inds=None):
    """`dets`: the REFERENCE's detections; `ref`: the oracle's.  Strict comparison (below) unless the reference dropped
    members of an exact plateau (see _plateau_mask; needs the oracle's logits `hm` [C,H,W] and the indices `more_inds` of
    `ref_more`): then every reference row must be one of the oracle's candidates, and every oracle row the reference lacks
    must sit on a plateau or in the group tied with the K-th score."""
    if hm is not None and not np.allclose(dets[:, 4], ref[:, 4], rtol=rtol, atol=1e-7):
        def find(r):                                  # candidate of the same class within 1e-4 of the row (no rounding keys)
            d = np.abs(ref_more - r[None, :])
            ok = (d[:, 5] == 0) & (d[:, :5].max(1) <= 1e-4 + 1e-5 * np.abs(r[:5]).max())
            return int(more_inds[np.argmax(ok)]) if ok.any() else None
        have = set()
        for r in dets:
            i = find(r)
            assert i is not None, "reference row %s is not an oracle peak" % (r,)
            have.add(i)
        plateau = _plateau_mask(hm).reshape(-1)
        kth = dets[-1, 4]
        dropped = 0
        for r, i in zip(ref_more, more_inds):
            if r[4] > kth * (1 + 1e-6) and int(i) not in have:
                assert plateau[int(i)], "oracle peak %s (index %d) is missing from the reference and is not on a plateau" % (r, i)
                dropped += 1
        assert 0 < dropped <= 8, dropped
        return
    _assert_dets_match_strict(dets, ref, ref_more, rtol, atol)


def _assert_dets_match_strict(dets, ref, ref_more=None, rtol=1e-5, atol=1e-5):
    """dets, ref: [K,6] sorted by score descending.  torch.topk leaves the order inside groups of equal scores
    unspecified, so rows are compared as multisets inside every group of equal score; for
    the group tied with the K-th score the membership itself is ambiguous, there each row must appear in `ref_more`
    (the reference decode with a larger K) when given."""
    K = ref.shape[0]
    np.testing.assert_allclose(dets[:, 4], ref[:, 4], rtol=rtol, atol=1e-7, err_msg="score lists differ")
    s = ref[:, 4]
    i = 0
    while i < K:
        j = i
        while j + 1 < K and abs(s[j + 1] - s[i]) <= 1e-9 * max(1.0, abs(s[i])):
            j += 1
        a = dets[i:j + 1]
        b = ref[i:j + 1] if (j + 1 < K or ref_more is None) else ref_more[np.abs(ref_more[:, 4] - s[i]) <= 1e-9]
        key = lambda r: tuple(np.round(r, 3))
        if j + 1 < K or ref_more is None:
            if j + 1 < K:
                assert sorted(map(key, a)) == sorted(map(key, b)), "rows differ in tie group %d..%d" % (i, j)
        else:
            bk = set(map(key, b))
            for r in a:
                assert key(r) in bk, "row %s not among the reference's tied candidates" % (r,)
        i = j + 1


def int8_mismatch(a, b):
    a, b = np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return int((a != b).sum())


def maxpool3s2_int8(a):
    """MaxPool2d(3, 2, 1) of an int8 grid [B,C,H,W] (the reference records the stem grid before its MaxPool)."""
    B, C, H, W = a.shape
    p = np.full((B, C, H + 2, W + 2), -128, np.int8)
    p[:, :, 1:-1, 1:-1] = a
    return np.max([p[:, :, i:i + H:2, j:j + W:2] for i in range(3) for j in range(3)], axis=0)


def check_reference_dets(g_dets, hm, wh, reg, K=100):
    """Reference detections `g_dets` [B,K,6] against the deterministic oracle decode of the (exact) head outputs
    hm (logits) / wh / reg [B,*,H,W]: strict up to torch.topk's tie order, plateau-aware where the reference's floating-point
    `hmax == heat` dropped one of two exactly equal neighbours."""
    from oracle import int_oracle as io
    hm, wh, reg = (np.asarray(a, np.float64) for a in (hm, wh, reg))
    odets, _ = io.ctdet_decode(hm, wh, reg, K)
    more, minds = io.ctdet_decode(hm, wh, reg, K + 60)
    for b in range(g_dets.shape[0]):
        assert_dets_match_tie_aware(g_dets[b], odets[b], more[b], hm=hm[b], more_inds=minds[b])


def assert_deform_f32_close(out, x, off, w, y_ref, stride, pad, dil, groups, dg, name=""):
    """fp32 deformable-conv output against the fp64 reference vector, EVERY element bounded (north star: 1e-4 relative).

    The fixture is the reference evaluated in fp64; the kernel sees the same tensors rounded to fp32.  Elements whose error
    exceeds 1e-4 * max(1, |y|max) are allowed only where (i) some tap of that output pixel samples within 2e-5 of an integer
    row / column -- there the fp32 coordinate may fall on the other side of floor(), which moves a bilinear corner by one
    pixel with a weight below 2e-5 -- and (ii) the error is still below the effect of that move, 4e-5 * sum|w| * max|x|."""
    from oracle import deform_ref
    f64 = np.float64
    x32, o32, w32 = (np.asarray(a, np.float32).astype(f64) for a in (x, off, w))
    y = deform_ref.deform_conv(x32, o32, w32, stride, pad, dil, groups, dg)          # the fp64 oracle on the fp32-rounded inputs
    scale = max(1.0, float(np.abs(y_ref).max()))
    np.testing.assert_allclose(y, y_ref, rtol=0, atol=2e-5 * scale, err_msg=name + ": oracle vs fixture")
    err = np.abs(np.asarray(out, f64) - y)
    bad = err > 1e-4 * scale
    if not bad.any():
        return 0
    B, _, Ho, Wo = y.shape
    kH, kW = w.shape[2], w.shape[3]
    hs = (np.arange(Ho) * stride - pad).reshape(1, Ho, 1)
    ws = (np.arange(Wo) * stride - pad).reshape(1, 1, Wo)
    near = np.zeros((B, Ho, Wo), bool)
    for g in range(dg):
        for i in range(kH):
            for j in range(kW):
                t = (g * kH * kW + i * kW + j) * 2
                for pos in (hs + i * dil + o32[:, t], ws + j * dil + o32[:, t + 1]):
                    near |= np.abs(pos - np.rint(pos)) < 2e-5
    assert not (bad & ~near[:, None]).any(), "%s: %d elements beyond 1e-4 away from any floor() boundary (max %.3g)" % (
        name, int((bad & ~near[:, None]).sum()), float(err[bad & ~near[:, None]].max()))
    cap = 4e-5 * float(np.abs(w32).reshape(w32.shape[0], -1).sum(1).max()) * float(np.abs(x32).max()) + 1e-4 * scale
    assert err[bad].max() <= cap, "%s: floor()-boundary outlier %.3g exceeds its bound %.3g" % (name, float(err[bad].max()), cap)
    return int(bad.sum())
