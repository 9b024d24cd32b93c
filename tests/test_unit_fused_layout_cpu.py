"""Host-side model of the index arithmetic of codenet_b200/csrc/unit_fused.cu / unit_s2_fused.cu / unit_fused.cuh (no GPU):
the `mid` tile's swizzle must be a bijection that keeps the first epilogue's stores and the stencil's loads off each other's
banks, the stencil must write exactly the K-major 128-byte-swizzled A tile of the second GEMM, the template-constant interleave
of uf_e2_fast must equal the plan's chunk table (= cat + channel_shuffle of the reference, shufflenetv2_dcn.py:29-34,102-114),
and the row arithmetic of the first GEMM's blocks must cover every halo pixel once."""
import numpy as np
import pytest

from codenet_b200.arch import NetConfig
from codenet_b200.plan import build_plan
from codenet_b200.synth import make_quant_state

UF_TW, UF_TH, UF_IW, UF_IH = 16, 8, 18, 10
US_IW, US_IH, US_MW = 33, 17, 34


def mid_swz(HP, c):
    return (c & 7) if HP == 128 else ((((c >> 1) & 1) << 2) | ((c >> 1) & 3))


def mid_off(HP, p, c, u):
    return (p * HP + (u << 4)) ^ (mid_swz(HP, c) << 4)


@pytest.mark.parametrize("HP,rowpix,rows,cols", [(64, UF_IW, UF_IH, UF_IW), (128, UF_IW, UF_IH, UF_IW), (64, US_MW, US_IH, US_IW)])
def test_mid_tile_swizzle_is_a_bijection(HP, rowpix, rows, cols):
    seen = set()
    for r in range(rows):
        for c in range(cols):
            p = r * rowpix + c
            for u in range(HP // 16):
                o = mid_off(HP, p, c, u)
                # a unit stays inside its pixel's 128-byte line (HP = 64: bit 2 of the swizzle swaps the two pixels of a line)
                assert o % 16 == 0 and 0 <= o < rows * rowpix * HP and o // 128 == (p * HP) // 128
                assert o not in seen
                seen.add(o)
    assert len(seen) == rows * cols * HP // 16


@pytest.mark.parametrize("HP", [64, 128])
def test_first_epilogue_stores_and_stencil_loads_are_conflict_free(HP):
    """E1: lane = consecutive A1 row = consecutive pixel, one 16-byte store per lane -> a quarter warp (8 lanes) must hit 8
    different 16-byte bank groups of the 128-byte line (row ends of the 18-pixel rows excepted).  S: a warp reads one 32-bit
    word of the same pixel per lane (HP = 128) or of two pixels 2 columns apart (HP = 64): 32 different banks."""
    bad = 0
    for p0 in range(0, 176, 8):
        groups = []
        for p in range(p0, p0 + 8):
            c = p % UF_IW
            groups.append((mid_off(HP, p, c, 1) % 128) // 16)
        wraps = (p0 % UF_IW) + 8 > UF_IW
        if len(set(groups)) != 8:
            assert wraps, (p0, groups)
            bad += 1
    assert bad <= 10
    CW = HP // 4
    for pg in range(8):
        for j in range(4):
            banks = set()
            for lane in range(32):
                cw = lane % CW
                c = 2 * (pg + lane // CW) + j if HP == 64 else 2 * pg + j
                if c >= UF_IW:
                    continue
                off = mid_off(HP, c, c, cw >> 2) + (cw & 3) * 4
                banks.add((off % 128) // 4)
            assert len(banks) == 32 or 2 * (pg + 1) + j >= UF_IW


def canonical_sw128(row, kbyte):
    return row * 128 + (((kbyte // 16) ^ (row % 8)) * 16) + kbyte % 16


@pytest.mark.parametrize("HP", [64, 128])
def test_stencil_writes_the_umma_a_tile(HP):
    CW = HP // 4
    RG = 256 // (CW * 8)
    RPG = UF_TH // RG
    seen = set()
    for tid in range(256):
        cw, pg, rg = tid % CW, (tid // CW) % 8, tid // (CW * 8)
        for j in range(2):
            m0 = rg * RPG * UF_TW + 2 * pg + j
            ao = m0 * 128 + (((cw >> 2) ^ (m0 & 7)) << 4) + (cw & 3) * 4
            for r in range(RPG):
                addr = ao + r * UF_TW * 128
                m = m0 + r * UF_TW
                assert addr == canonical_sw128(m, cw * 4)
                seen.add((m, cw))
    assert len(seen) == 128 * CW                                           # every output pixel x channel word once


def test_first_gemm_blocks_cover_the_halo_tile():
    """stride 1: rows 0..127 (block 0) and 64..191 (block 1, rows >= 128 new) cover the 180 halo pixels once; the row / 18
    multiply-shift is exact.  stride 2: five blocks cover 561 rows, row / 33 likewise."""
    done = set()
    for blk in range(2):
        for q in range(4):
            for lane in range(32):
                row = blk * 64 + q * 32 + lane
                valid = blk == 0 or (128 <= row < 180)
                assert ((row * 3641) >> 16) == row // 18
                if valid:
                    assert row not in done
                    done.add(row)
    assert done == set(range(180))
    for row in range(640):
        assert ((row * 1986) >> 16) == row // 33
    assert 4 * 128 + 49 == US_IW * US_IH


def _interleave_fast(HP, PG, new_cols, pass_bytes):
    """uf_e2_fast for one pixel: new_cols[N3] (int8 per GEMM column), pass_bytes[128] -> the 2*HP output bytes."""
    Gp = (PG + 7) & ~7
    out = np.zeros(2 * HP, np.int8)
    for G in range(2):
        for j in range(HP // 16):
            cnt = min(max(PG - 8 * j, 0), 8)
            col, pass_off, dst = G * Gp + 8 * j, G * PG + 8 * j, G * HP + 16 * j
            o = np.zeros(16, np.int8)
            for i in range(cnt):
                o[2 * i] = pass_bytes[pass_off + i]
                o[2 * i + 1] = new_cols[col + i]
            out[dst:dst + 16] = o
    return out


@pytest.mark.parametrize("w2,stage,HP,PG", [(False, 1, 64, 29), (False, 2, 128, 58), (True, 1, 128, 61)])
def test_template_interleave_equals_the_plans_chunk_table(calib, w2, stage, HP, PG, golden):
    """The fast kernels hard-wire the chunk table as a function of (HP, PG); the plan's table for the same unit must be exactly
    that function, and applying it must give cat + channel_shuffle: out[2k] = pass[k], out[2k+1] = new[k] in the half layout."""
    cfg = NetConfig(num_classes=20, w2=w2, maxpool=w2)
    cal = golden("codenet_w2mp_calib.npz") if w2 else calib
    st = make_quant_state(cfg, cal, "round", 256)
    plan = build_plan(cfg, st, 256, 256, "round")
    op = next(o for o in plan.ops if o.name == "layer%d.1.pw3" % stage)
    chunks = [tuple(int(v) for v in c) for c in op.a["chunks"]]
    Gp = (PG + 7) & ~7
    want = []
    for G in range(2):
        for j in range(HP // 16):
            cnt = min(max(PG - 8 * j, 0), 8)
            want.append((G * Gp + 8 * j, cnt, G * PG + 8 * j, G * HP + 16 * j) if cnt else (0, 0, -1, G * HP + 16 * j))
    assert sorted(chunks, key=lambda c: c[3]) == sorted(want, key=lambda c: c[3])
    rng = np.random.default_rng(1)
    h = 2 * PG
    new_logical = rng.integers(-128, 128, h).astype(np.int8)
    pass_logical = rng.integers(-128, 128, h).astype(np.int8)
    N3 = op.a["N"]
    new_cols = np.zeros(N3, np.int8)
    col = np.where(np.arange(h) < PG, np.arange(h), Gp + np.arange(h) - PG)       # plan.py emit_pw: GEMM column of branch channel k
    new_cols[col] = new_logical
    pass_bytes = np.zeros(128, np.int8)
    pass_bytes[:h] = pass_logical                                                # first half of the stage tensor (bytes [0, h))
    out = _interleave_fast(HP, PG, new_cols, pass_bytes)
    shuffled = np.empty(2 * h, np.int8)                                          # channel_shuffle(cat(pass, new), 2)
    shuffled[0::2], shuffled[1::2] = pass_logical, new_logical
    tout = plan.tensors[op.a["out_t"]]
    np.testing.assert_array_equal(out[tout.phys(np.arange(2 * h))], shuffled)
    pad = np.ones(2 * HP, bool)
    pad[tout.phys(np.arange(2 * h))] = False
    assert not out[pad].any()                                                    # pad bytes of the pixel stay zero
