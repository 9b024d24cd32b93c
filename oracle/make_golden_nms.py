"""Generates tests/golden/soft_nms_kat.npz from the reference's own compiled Cython `soft_nms`
(lib/models/external/nms.pyx:77-170, built unmodified by oracle/build_ref.build_nms; container only).
TEST INFRASTRUCTURE.  Cases: random boxes per class as merge_outputs sees them (lib/detectors/ctdet.py:59-66), the three
methods (0 hard, 1 linear, 2 gaussian = the one CoDeNet uses with Nt=0.5), low-score sets that trigger the discard swap."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402


def main():
    build_ref.build_nms()
    ref = build_ref.load_nms()
    rng = np.random.default_rng(7)
    out = {}
    for t in range(12):
        n = int(rng.integers(1, 120))
        c = rng.uniform(0, 300, (n, 2)); wh = rng.uniform(4, 120, (n, 2))
        b = np.concatenate([c - wh / 2, c + wh / 2, rng.uniform(0, 1, (n, 1)) ** 2], 1).astype(np.float32)
        if t % 3 == 0:
            b[:, 4] *= 0.01                       # many scores end below the 0.001 threshold
        if t % 4 == 1:
            b = np.concatenate([b, b[: n // 2] + np.float32(0.5)], 0)       # near-duplicates (two scales of one image)
        for method in (0, 1, 2):
            a = b.copy()
            keep = ref.soft_nms(a, Nt=0.5, method=method)
            out["c%d_m%d_in" % (t, method)] = b
            out["c%d_m%d_out" % (t, method)] = a
            out["c%d_m%d_keep" % (t, method)] = np.asarray(len(keep), np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "soft_nms_kat.npz"), **out)
    print("wrote soft_nms_kat.npz:", len(out) // 3, "cases")


if __name__ == "__main__":
    main()
