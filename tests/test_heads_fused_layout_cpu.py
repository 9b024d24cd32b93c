"""Host-side model of the index arithmetic of codenet_b200/csrc/heads_fused.cu and dw_tma.cu (no GPU): the shared-memory
A tile the depthwise threads write must be exactly the K-major 128-byte-swizzled operand the UMMA descriptor describes, the
weight loader must produce the same layout for B, and the tile walk must cover every output pixel once."""
import numpy as np
import pytest

HF_TW, HF_TH = 4, 8


def canonical_sw128(row, kbyte, rows_per_block):
    """Byte offset of element (row, kbyte) of a K-major operand with 128-byte rows, 8-row groups 1024 bytes apart and the
    128B swizzle (16-byte unit index XOR row % 8) -- what make_smem_desc() + TMA SWIZZLE_128B mean in pw_gemm.cu."""
    kb, kk = kbyte // 128, kbyte % 128
    return kb * rows_per_block * 128 + row * 128 + (((kk // 16) ^ (row % 8)) * 16) + kk % 16


@pytest.mark.parametrize("pitch", [128, 160, 192])
def test_a_tile_written_once_in_umma_layout(pitch):
    cw_total = pitch // 4
    seen = {}
    for tid in range(cw_total * HF_TW):
        cw, pc = tid % cw_total, tid // cw_total
        kbyte = cw * 4
        a_unit = (kbyte & 127) >> 4
        a_base = (kbyte >> 7) * 16384 + (kbyte & 15)
        for r in range(HF_TH):
            for yp in range(2):
                m0 = (2 * r + yp) * 8 + 2 * pc
                for m in (m0, m0 + 1):
                    addr = a_base + m * 128 + ((a_unit ^ (m & 7)) << 4)
                    assert addr == canonical_sw128(m, kbyte, 128)
                    assert addr % 4 == 0 and addr + 4 <= 2 * 16384
                    assert (m, kbyte) not in seen
                    seen[m, kbyte] = addr
    assert len(seen) == 128 * cw_total                       # every (pixel row, channel word) of the 128 x K tile
    assert len(set(seen.values())) == len(seen)              # no two words share an address


def test_a_tile_rows_are_the_tile_pixels():
    """Row m of the UMMA = output pixel (oy, ox) = (m >> 3, m & 7) of the 8 x 16 tile: the epilogue's mapping."""
    rows = set()
    for pc in range(HF_TW):
        for r in range(HF_TH):
            for yp in range(2):
                for xp in range(2):
                    m = (2 * r + yp) * 8 + 2 * pc + xp
                    assert (m >> 3, m & 7) == (2 * r + yp, 2 * pc + xp)
                    rows.add(m)
    assert rows == set(range(128))


@pytest.mark.parametrize("NB,Kp", [(16, 256), (32, 256), (96, 256), (32, 128)])
def test_weight_loader_layout(NB, Kp):
    upr = Kp >> 4
    seen = set()
    for i in range(NB * upr):
        n, c = i // upr, i % upr
        kb, cc = c >> 3, c & 7
        dst = kb * NB * 128 + n * 128 + ((cc ^ (n & 7)) << 4)
        assert dst == canonical_sw128(n, 16 * c, NB)
        assert dst not in seen
        seen.add(dst)
    assert max(seen) + 16 <= 2 * NB * 128


@pytest.mark.parametrize("Hs,Ws,batch", [(64, 64, 3), (32, 32, 2), (40, 40, 1)])
def test_tile_walk_covers_every_output_pixel_once(Hs, Ws, batch):
    tiles_x, tiles_y = Ws // HF_TW, Hs // HF_TH
    cover = np.zeros((batch, 2 * Hs, 2 * Ws), np.int32)
    for tile in range(batch * tiles_x * tiles_y):
        tx = tile % tiles_x
        ty = (tile // tiles_x) % tiles_y
        b = tile // (tiles_x * tiles_y)
        for m in range(128):
            cover[b, ty * 2 * HF_TH + (m >> 3), tx * 2 * HF_TW + (m & 7)] += 1
    assert (cover == 1).all()


@pytest.mark.parametrize("pitch,stride", [(64, 1), (128, 1), (256, 1), (32, 2), (64, 2), (128, 2), (256, 2)])
def test_dw_tma_tile_geometry(pitch, stride):
    """dw_tma.cu: 128 threads = channel words x pixel pairs (stride 1) / pixels (stride 2) of one tile row; every staged pixel a
    thread reads lies inside the TMA box."""
    TW = (512 if stride == 2 else 1024) // pitch
    TH = 8
    cw_total = pitch // 4
    per_row = TW // 2 if stride == 1 else TW
    assert cw_total * per_row == 128
    in_w = 2 * TW + 1 if stride == 2 else TW + 2
    in_h = 2 * TH + 1 if stride == 2 else TH + 2
    for pg in range(per_row):
        cols = [2 * pg + j for j in range(3 if stride == 2 else 4)]
        assert max(cols) < in_w
        out_cols = [pg] if stride == 2 else [2 * pg, 2 * pg + 1]
        assert max(out_cols) < TW
    rows_read = 1 + 2 * TH if stride == 2 else 2 + TH
    assert rows_read == in_h
    assert in_w <= 256 and in_h <= 256 and (in_h * in_w * pitch) > 0
