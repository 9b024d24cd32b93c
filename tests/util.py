"""Helpers shared by the parity tests."""
import numpy as np


def assert_dets_match_tie_aware(dets, ref, ref_more=None, rtol=1e-5, atol=1e-5):
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
