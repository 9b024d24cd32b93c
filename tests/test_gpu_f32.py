"""Float (fp32) model path on the GPU against the reference evaluated in fp64 (tolerance 1e-4 relative, north star)."""
import numpy as np
import pytest

from codenet_b200.arch import NetConfig
from codenet_b200.synth import make_raw_state, make_images, state_digest
from oracle import int_oracle as io

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gemm", ["fp32", "tf32x3"])
@pytest.mark.parametrize("tag,cfg", [("1x", NetConfig(num_classes=20)), ("2x_coco", NetConfig(num_classes=80, w2=True))])
def test_float_model_matches_reference_fp64(golden, tag, cfg, gemm):
    import torch
    from codenet_b200.engine_f32 import EngineF32
    g = golden("codenet_float_%s_256.npz" % tag)
    raw = make_raw_state(cfg, 0)
    assert state_digest(raw) == str(g["digest"])
    for k in g.files:
        if k.startswith("bn/"):
            raw[k[3:]] = g[k]
    eng = EngineF32(cfg, raw, gemm=gemm)
    x = torch.from_numpy(make_images(2, 256, seed=2)[:1].copy()).cuda()
    dets, inds, v = eng.detect(x)
    torch.cuda.synchronize()
    # Tolerance: 1e-4 relative (L2) against the fp64 reference -- or the deviation of the reference's OWN fp32 evaluation
    # from its fp64 one where that is larger (the random synthetic network is ill-conditioned: up to 1.4e-3 absolute on
    # values ~4, 2.7e-4 L2 on the small `reg` map); both yardsticks are recorded in the golden file by the generating script,
    # and our worst element must not exceed the reference's worst fp32 element.  gemm="tf32x3" (1x1 convs on the tensor cores,
    # products from a 3-way TF32 split, pw_tf32.cu) is the throughput mode: 1e-4 or 1.5x the reference's own fp32 deviation
    # (measured: hm 4.1e-5, wh 4.1e-5, reg 1.9e-4 on the 2x model = at the reference's own level; reg 3.1e-4 against 2.7e-4 on
    # the 1x model), and its worst element may be up to twice the reference's own worst fp32 element (measured 1.4x).
    for name, key in (("hm", "hm_logit"), ("wh", "wh"), ("reg", "reg")):
        got, ref = v[name].cpu().numpy().astype(np.float64), g[key].astype(np.float64)
        l2 = np.sqrt(((got - ref) ** 2).sum() / (ref ** 2).sum())
        assert l2 <= max(1e-4, (1.5 if gemm == "tf32x3" else 1.0) * float(g["ref_fp32_l2rel/" + name])), (name, l2, float(g["ref_fp32_l2rel/" + name]))
        assert np.abs(got - ref).max() <= (2.0 if gemm == "tf32x3" else 1.0) * float(g["ref_fp32_maxerr/" + name]) + 1e-6, \
            (name, float(np.abs(got - ref).max()))
    # detections: decode the engine's own heads with the oracle (exact indices), and agree with the reference's boxes
    heads = {k: t.cpu().numpy().astype(np.float64) for k, t in v.items()}
    odets, oinds = io.ctdet_decode(heads["hm"], heads["wh"], heads["reg"], 100)
    np.testing.assert_array_equal(inds.cpu().numpy(), oinds)
    d = dets.cpu().numpy()
    np.testing.assert_allclose(d, odets, rtol=1e-5, atol=1e-4)
    ref = g["dets"][0]
    # same detections as the reference up to score ties / near-ties: compare the score-sorted score list and the best box
    np.testing.assert_allclose(np.sort(d[0, :, 4])[::-1][:50], np.sort(ref[:, 4])[::-1][:50], rtol=1e-3)
    np.testing.assert_allclose(d[0, 0, :4], ref[0, :4], rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("C,Co,ppi,B,in_ct,in_off,out_ct,out_off,out_cs,relu", [
    (24, 122, 256, 3, 24, 0, 244, 0, 2, 1),            # layer1.0 pw5: even channels of the shuffled output
    (122, 122, 512, 2, 244, 122, 244, 1, 2, 1),        # stride-1 unit: second half in, odd channels out
    (244, 244, 256, 5, 244, 0, 244, 0, 1, 1),          # two N tiles of 122 channels
    (976, 300, 256, 2, 976, 0, 300, 0, 1, 0),          # long K, three N tiles of 100 (BN = 112)
    (64, 80, 1024, 1, 64, 0, 84, 0, 1, 0),             # hm.out into the shared heads tensor
    (64, 2, 256, 2, 64, 0, 84, 80, 1, 0),              # wh.out: N = 16 with two real channels
    (2153, 256, 256, 1, 2153, 0, 256, 0, 1, 1),        # up0.channel: K not a multiple of 16
])
def test_pw_tf32x3_matches_fp64(C, Co, ppi, B, in_ct, in_off, out_ct, out_off, out_cs, relu):
    """cdn_pw_slice_tf32x3 against the fp64 product: every element within 2^-19 of sum |w||x| (the 3-way TF32 split keeps each
    product to ~2^-21, the fp32 accumulator adds its roundings), untouched output channels stay untouched, and the result
    equals the SIMT fp32 kernel's to 1e-5 relative L2."""
    import torch
    from codenet_b200 import _lib
    from gpu_util import ptr, stream
    L = _lib.load()
    rng = np.random.default_rng(C * 7 + Co)
    x = rng.normal(0, 1, (B, in_ct, ppi)).astype(np.float32)
    w = (rng.normal(0, 1, (Co, C)) / np.sqrt(C)).astype(np.float32)
    bias = rng.normal(0, 1, Co).astype(np.float32)
    tx, tw, tb = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(bias).cuda()
    n = int(L.cdn_pw_tf32x3_packed_floats(Co, C))
    packed = torch.empty(n, device="cuda")
    _lib.check(L.cdn_pw_tf32x3_pack(ptr(tw), Co, C, ptr(packed), stream()))
    out = torch.full((B, out_ct, ppi), -77.0, device="cuda")
    _lib.check(L.cdn_pw_slice_tf32x3(ptr(tx), in_ct, in_off, C, ptr(packed), ptr(tb), ptr(out), out_ct, out_off, out_cs, Co, relu,
                                     B, ppi, stream()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    xs = x[:, in_off:in_off + C].astype(np.float64)
    ref = np.einsum("oc,bcp->bop", w.astype(np.float64), xs) + bias[None, :, None]
    mag = np.einsum("oc,bcp->bop", np.abs(w).astype(np.float64), np.abs(xs)) + np.abs(bias)[None, :, None]
    if relu:
        ref = np.maximum(ref, 0)
    ch = out_off + np.arange(Co) * out_cs
    err = np.abs(got[:, ch] - ref)
    assert (err <= mag * 2.0 ** -19).all(), float((err / mag).max())
    rest = np.setdiff1d(np.arange(out_ct), ch)
    assert (got[:, rest] == -77.0).all()
    out2 = torch.full((B, out_ct, ppi), -77.0, device="cuda")
    _lib.check(L.cdn_pw_slice_f32(ptr(tx), in_ct, in_off, C, ptr(tw), ptr(tb), ptr(out2), out_ct, out_off, out_cs, Co, relu, B, ppi, stream()))
    d = (out2.cpu().numpy()[:, ch].astype(np.float64) - got[:, ch])
    assert np.sqrt((d ** 2).sum() / (ref ** 2).sum()) <= 1e-5
    # pixel counts that are not a multiple of 256 are refused (the caller keeps cdn_pw_slice_f32 for them)
    assert L.cdn_pw_slice_tf32x3(ptr(tx), in_ct, in_off, C, ptr(packed), ptr(tb), ptr(out), out_ct, out_off, out_cs, Co, relu,
                                 B, 128, stream()) == -1


@pytest.mark.parametrize("B,C,H,W,stride", [(2, 5, 16, 16, 1), (1, 7, 32, 32, 1), (2, 3, 64, 64, 1), (1, 4, 128, 128, 1), (1, 3, 8, 256, 1),
                                            (2, 4, 10, 12, 1), (1, 3, 9, 7, 1), (2, 4, 16, 16, 2), (1, 3, 9, 7, 2)])
def test_dw3x3_f32_matches_numpy(B, C, H, W, stride):
    """cdn_dw3x3_f32 (depthwise 3x3, pad 1, + bias + ReLU: the float BaseNode's second conv, shufflenetv2_dcn.py:77-80): every
    kernel variant (4 x 4 outputs per thread with shuffled halos, with loaded halos, 4 x 1, scalar) against the fp64 sum."""
    import torch
    from codenet_b200 import _lib
    from gpu_util import ptr, stream
    rng = np.random.default_rng(B * 100 + H + W)
    x = rng.normal(0, 1, (B, C, H, W)).astype(np.float32)
    w = rng.normal(0, 1, (C, 3, 3)).astype(np.float32)
    bias = rng.normal(0, 1, C).astype(np.float32)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1)))
    ref = np.zeros((B, C, Ho, Wo))
    for i in range(3):
        for j in range(3):
            ref += w[None, :, i, j, None, None] * xp[:, :, i:i + stride * Ho:stride, j:j + stride * Wo:stride][:, :, :Ho, :Wo]
    ref = np.maximum(ref + bias[None, :, None, None], 0)
    tx, tw, tb = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(bias).cuda()
    out = torch.full((B, C, Ho, Wo), -7.0, device="cuda")
    _lib.check(_lib.load().cdn_dw3x3_f32(ptr(tx), ptr(tw), ptr(tb), ptr(out), B, C, H, W, stride, 1, stream()))
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("ppi", [256, 6])
def test_copy_channels_f32(ppi):
    """cdn_copy_channels_f32 = the pass-through half of a stride-1 unit written to the even channels (cat + channel_shuffle,
    shufflenetv2_dcn.py:29-34,108-114); 16-byte and scalar forms."""
    import torch
    from codenet_b200 import _lib
    from gpu_util import ptr, stream
    B, ct, n = 3, 10, 5
    x = torch.randn(B, ct, ppi, device="cuda")
    out = torch.full((B, ct, ppi), -7.0, device="cuda")
    _lib.check(_lib.load().cdn_copy_channels_f32(ptr(x), ct, 0, ptr(out), ct, 0, 2, n, B, ppi, stream()))
    torch.cuda.synchronize()
    assert torch.equal(out[:, 0::2], x[:, :n]) and bool((out[:, 1::2] == -7.0).all())


@pytest.mark.parametrize("B,C,h,w", [(2, 5, 8, 8), (1, 3, 64, 64), (2, 4, 3, 12)])
def test_dw3x3_over_virtual_upsample_equals_upsample_then_dw(B, C, h, w):
    """cdn_dw3x3_up2_f32(a) == cdn_dw3x3_f32(cdn_upsample2x_f32(a)) bit for bit (same taps in the same order): the heads' fused
    form of nearest x2 upsampling followed by their depthwise conv."""
    import torch
    from codenet_b200 import _lib
    from gpu_util import ptr, stream
    L = _lib.load()
    a = torch.randn(B, C, h, w, device="cuda")
    wt, bias = torch.randn(C, 3, 3, device="cuda"), torch.randn(C, device="cuda")
    up = torch.empty(B, C, 2 * h, 2 * w, device="cuda")
    _lib.check(L.cdn_upsample2x_f32(ptr(a), ptr(up), B * C, h, w, stream()))
    assert torch.equal(up, a.repeat_interleave(2, 2).repeat_interleave(2, 3))
    want, got = torch.empty_like(up), torch.empty_like(up)
    _lib.check(L.cdn_dw3x3_f32(ptr(up), ptr(wt), ptr(bias), ptr(want), B, C, 2 * h, 2 * w, 1, 1, stream()))
    _lib.check(L.cdn_dw3x3_up2_f32(ptr(a), ptr(wt), ptr(bias), ptr(got), B, C, h, w, 1, stream()))
    torch.cuda.synchronize()
    assert torch.equal(got, want)


@pytest.mark.parametrize("B,Co,H,W,stride", [(2, 24, 64, 64, 4), (1, 24, 32, 48, 4), (2, 24, 30, 26, 4), (1, 24, 32, 32, 2), (1, 7, 16, 20, 4)])
def test_stem_conv3x3_f32_matches_numpy(B, Co, H, W, stride):
    """cdn_conv3x3_f32 = the float model's stem conv (3 -> Co, 3x3, pad 1, stride 4 or 2, + folded BN bias + ReLU,
    shufflenetv2_dcn.py:196-203): the stride-4 kernel with 16-byte loads and the general scalar kernel against the fp64 sum."""
    import torch
    from codenet_b200 import _lib
    from gpu_util import ptr, stream
    rng = np.random.default_rng(H * 7 + W + Co)
    x = rng.normal(0, 1, (B, 3, H, W)).astype(np.float32)
    w = (rng.normal(0, 1, (Co, 3, 3, 3)) / 5).astype(np.float32)
    bias = rng.normal(0, 1, Co).astype(np.float32)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1)))
    ref = np.zeros((B, Co, Ho, Wo))
    for i in range(3):
        for j in range(3):
            ref += np.einsum("oc,bchw->bohw", w[:, :, i, j].astype(np.float64), xp[:, :, i:i + stride * Ho:stride, j:j + stride * Wo:stride][:, :, :Ho, :Wo])
    ref = np.maximum(ref + bias[None, :, None, None], 0)
    tx, tw, tb = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(bias).cuda()
    out = torch.full((B, Co, Ho, Wo), -7.0, device="cuda")
    _lib.check(_lib.load().cdn_conv3x3_f32(ptr(tx), ptr(tw), ptr(tb), ptr(out), B, 3, Co, H, W, stride, 1, stream()))
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=3e-6, atol=3e-6)


@pytest.mark.parametrize("B,C,h,w,bound", [(2, 19, 8, 8, 8), (1, 40, 16, 12, 4), (2, 5, 3, 5, 8)])
def test_deform_module_over_virtual_upsample_equals_upsample_then_module(B, C, h, w, bound):
    """cdn_deform_dw_up2_f32_ws(a) == cdn_deform_dw_f32_ws(cdn_upsample2x_f32(a)) bit for bit: the up path's
    Upsample -> DeformConvWithOffsetScaleBoundPositive (shufflenetv2_dcn.py:286-300) without the upsampled tensor."""
    import ctypes as C_
    import torch
    from codenet_b200 import _lib
    from gpu_util import ptr, stream
    L = _lib.load()
    a = torch.randn(B, C, h, w, device="cuda")
    ws, wd = torch.randn(C, device="cuda") * 0.5, torch.randn(C, 3, 3, device="cuda")
    up = torch.empty(B, C, 2 * h, 2 * w, device="cuda")
    _lib.check(L.cdn_upsample2x_f32(ptr(a), ptr(up), B * C, h, w, stream()))
    need = int(L.cdn_deform_dw_f32_ws_bytes(B, 2 * h, 2 * w, 1))
    wsp = torch.empty(need, dtype=torch.uint8, device="cuda")
    want, got = torch.empty_like(up), torch.empty_like(up)
    _lib.check(L.cdn_deform_dw_f32_ws(ptr(up), ptr(ws), C_.c_float(1.3), bound, ptr(wd), ptr(want), B, C, 2 * h, 2 * w, 1, ptr(wsp), need, stream()))
    _lib.check(L.cdn_deform_dw_up2_f32_ws(ptr(a), ptr(ws), C_.c_float(1.3), bound, ptr(wd), ptr(got), B, C, h, w, ptr(wsp), need, stream()))
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    assert float(want.abs().max()) > 0


def test_float_model_gemm_modes_agree_on_a_non_square_batch():
    """384 x 640 input, batch 3 (pixel counts 15360 / 3840 on the tensor-core path, 960 / 240 on the SIMT fallback, tile counts below
    and above the SM count): the two 1x1-conv implementations must give the same network within the float-path bound, and the
    same detections up to near-ties."""
    import torch
    from codenet_b200.engine_f32 import EngineF32
    cfg = NetConfig(num_classes=20)
    raw = make_raw_state(cfg, 0)
    rng = np.random.default_rng(11)
    x = torch.from_numpy(rng.normal(0, 1, (3, 3, 384, 640)).astype(np.float32)).cuda()
    outs = {}
    for gemm in ("fp32", "tf32x3"):
        eng = EngineF32(cfg, raw, gemm=gemm)
        dets, inds, v = eng.detect(x)
        torch.cuda.synchronize()
        outs[gemm] = (dets.cpu().numpy(), {k: t.cpu().numpy().astype(np.float64) for k, t in v.items()})
    for name in ("hm", "wh", "reg"):
        a, b = outs["fp32"][1][name], outs["tf32x3"][1][name]
        assert a.shape == (3, {"hm": 20, "wh": 2, "reg": 2}[name], 96, 160)
        l2 = np.sqrt(((a - b) ** 2).sum() / (a ** 2).sum())
        assert l2 <= 5e-4, (name, l2)
    np.testing.assert_allclose(np.sort(outs["tf32x3"][0][..., 4], 1)[:, ::-1][:, :30], np.sort(outs["fp32"][0][..., 4], 1)[:, ::-1][:, :30], rtol=2e-3)


@pytest.mark.parametrize("Co,C", [(122, 24), (244, 61), (80, 64), (2, 64), (300, 130)])
def test_tf32x3_weight_packing_equals_the_restatement(Co, C):
    """cdn_pw_tf32x3_pack against oracle/tf32_split.pack_weights, bit for bit (rounding to TF32 with ties away from zero, exact
    remainder, N-tile / padding layout)."""
    import torch
    from codenet_b200 import _lib
    from gpu_util import ptr, stream
    from oracle import tf32_split as ts
    rng = np.random.default_rng(Co + C)
    w = rng.normal(0, 1, (Co, C)).astype(np.float32)
    w[0, 0], w[-1, -1] = np.float32(1 + 2 ** -11), np.float32(-(1 + 2 ** -11))         # exact ties
    L = _lib.load()
    n = int(L.cdn_pw_tf32x3_packed_floats(Co, C))
    want = ts.pack_weights(w)
    assert want.size == n
    packed = torch.empty(n, device="cuda")
    _lib.check(L.cdn_pw_tf32x3_pack(ptr(torch.from_numpy(w).cuda()), Co, C, ptr(packed), stream()))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(packed.cpu().numpy().view(np.uint32), want.view(np.uint32))
