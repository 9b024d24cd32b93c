"""Exactness of the integer requantisation the int8 kernels evaluate (host solver, no GPU needed).

The library replaces  q = clamp(rint(fl64(fl64(v*M) + B)), lo, 127)  -- the per-channel formula of DESIGN.md, i.e. the
reference's `rint(s_out*y - z_out)` with y = acc/(sigma_c*s_x) + b'_c (portable_quantizer/quantization_utils/
quant_utils.py:31-39,58-73; quant_modules.py:364-419) -- by  q = sat8(max(((v*Mi + Bi) >> 32) >> sh, lo))  with (Mi, sh, Bi)
from cdn_rq_int_solve.  The claim is equality for EVERY integer accumulator of the layer's range; it is checked here
exhaustively (all v) against a numpy evaluation of the fp64 formula."""
import ctypes as C

import numpy as np
import pytest

from codenet_b200 import _lib


def solve(M, B, lo, vmin, vmax):
    lib = _lib.load()
    Mi, sh, Bi = C.c_int32(), C.c_int32(), C.c_int64()
    r = lib.cdn_rq_int_solve(float(M), float(B), int(lo), int(vmin), int(vmax), C.byref(Mi), C.byref(sh), C.byref(Bi))
    return r, Mi.value, sh.value, Bi.value


def ref_q(v, M, B, lo):
    t = (v.astype(np.float64) * np.float64(M)) + np.float64(B)          # two separately rounded fp64 operations
    return np.clip(np.rint(t), lo, 127).astype(np.int64)


def int_q(v, Mi, sh, Bi, lo):
    x = v.astype(np.int64) * np.int64(Mi) + np.int64(Bi)                # exact: the solver bounds |x| < 2^62
    return np.clip(np.maximum((x >> 32) >> sh, lo), -128, 127)


def check_all(M, B, lo, vmax, chunk=1 << 22, may_refuse=False):
    r, Mi, sh, Bi = solve(M, B, lo, -vmax, vmax)
    if may_refuse and r != 0:
        return False
    assert r == 0, (M, B, lo, vmax)
    assert (0 < Mi < 2 ** 31 or M == 0.0) and 0 <= sh < 32
    for a in range(-vmax, vmax + 1, chunk):
        v = np.arange(a, min(a + chunk, vmax + 1), dtype=np.int64)
        bad = np.nonzero(ref_q(v, M, B, lo) != int_q(v, Mi, sh, Bi, lo))[0]
        assert bad.size == 0, (M, B, lo, vmax, int(v[bad[0]]))
    return True


def test_exhaustive_random_channels():
    rng = np.random.default_rng(11)
    for _ in range(60):
        vmax = int(rng.choice([255 * 9 * 7, 255 * 58 * 5, 255 * 464 * 4, 255 * 1024 * 3]))
        M = 255.0 / (vmax * rng.uniform(0.05, 1.5))                     # output range covers 5 % .. 150 % of the accumulator's
        B = rng.uniform(-300, 300)
        lo = int(rng.choice([-128, -128, -101, 0]))
        check_all(M, B, lo, vmax)


def test_exhaustive_full_magic_range():
    # the widest accumulator the GEMM accepts (|v| < 2^22) with the smallest multipliers that implies
    rng = np.random.default_rng(12)
    for _ in range(4):
        check_all(255.0 / (2 ** 22 * rng.uniform(0.3, 1.0)), rng.uniform(-200, 200), -128, 2 ** 22 - 1)


def test_adversarial_ties_and_powers_of_two():
    # Exact .5 ties at integer accumulators: round-half-even makes the step spacing alternate (7, 9, 7, 9 for M = 1/8),
    # which no single line reproduces -- the solver must REFUSE those (the layer then keeps the guarded fp32 sequence)
    # and must never return a pair that is wrong anywhere.  Without exact ties it must succeed.
    solved = refused = 0
    for M in (2.0 ** -3, 2.0 ** -7, 3 * 2.0 ** -9, 0.2499999, 0.1):
        for B in (0.0, 0.5, -0.5, 17.5, -128.5, 126.5, 1e-9, -1e-9):
            for lo in (-128, -77):
                if check_all(M, B, lo, 5000, may_refuse=True):
                    solved += 1
                else:
                    refused += 1
    assert refused > 0 and solved > 0
    assert solve(2.0 ** -3, 0.0, -128, -5000, 5000)[0] != 0
    for B in (0.3123, 0.7701, -17.2345):                # no accumulator lands on a tie: must solve
        check_all(0.1, B, -128, 5000)
        check_all(0.2499999, B, -128, 5000)


def test_degenerate_ranges():
    check_all(1e-6, 3.2, -128, 1000)              # whole range maps to one level
    check_all(1e-6, 500.0, -128, 1000)            # always saturated high
    check_all(1e-6, -500.0, -128, 1000)           # always saturated low
    check_all(0.01, 0.0, 127, 100000)             # lo == 127: nothing to solve
    check_all(0.0, 0.0, -128, 1000)               # padding channel (M == 0): constant rint(B)
    check_all(0.0, -3.5, -128, 1000)
    check_all(0.0, 2.5, -1, 1000)


def test_refuses_what_it_cannot_represent():
    assert solve(0.75, 0.0, -128, -1000, 1000)[0] != 0          # multiplier >= 0.5
    assert solve(-0.1, 0.0, -128, -1000, 1000)[0] != 0          # non-positive multiplier
    assert solve(float("nan"), 0.0, -128, -1000, 1000)[0] != 0
