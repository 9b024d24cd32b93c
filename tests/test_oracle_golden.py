"""The oracle is only trusted after it reproduces vectors made by the UNMODIFIED reference (oracle/make_golden.py)."""
import numpy as np
import pytest

from codenet_b200.arch import NetConfig
from codenet_b200.synth import make_quant_state, make_images
from oracle import int_oracle as io
from oracle import deform_ref
from util import check_reference_dets, int8_mismatch

CFG = NetConfig(num_classes=20)


def test_activation_quant_kat(golden):
    g = golden("quant_kat.npz")
    for i in range(3):
        lo, hi = g["act%d_range" % i]
        s, z = io.act_params(lo, hi)
        q = np.rint(s * g["act%d_x" % i] - z)
        np.testing.assert_array_equal((q + z) / s, g["act%d_y" % i])
    # stateful init of QuantAct = batch min/max (quant_modules.py:209-212)
    assert g["qact_min"][0] == g["qact_x"].min() and g["qact_max"][0] == g["qact_x"].max()
    s, z = io.act_params(g["qact_min"][0], g["qact_max"][0])
    np.testing.assert_array_equal((np.rint(s * g["qact_x"] - z) + z) / s, g["qact_y"])


@pytest.mark.parametrize("bits", [4, 8])
def test_weight_quant_kat(golden, bits):
    g = golden("quant_kat.npz")
    wq, sigma = io.quant_weight(g["w%d_x" % bits], bits)
    np.testing.assert_array_equal(wq / sigma.reshape(-1, 1, 1, 1), g["w%d_y" % bits])
    assert wq.min() >= -(2 ** (bits - 1)) and wq.max() <= 2 ** (bits - 1) - 1


@pytest.mark.parametrize("tag", ["pct_wide", "pct_narrow", "pct_dw"])
def test_weight_percentile_kat(golden, tag):
    """--wt-percentile (quant_modules.py:296-309, :382-395): the oracle's and the plan compiler's weight ranges against the
    reference's own Quant_Conv2d(weight_percentile=True) -- 1200 inputs per row (k-th values), 5 and 9 (0.95 * min / max)."""
    from codenet_b200 import plan
    g = golden("quant_kat.npz")
    w, x = g[tag + "_w"], g[tag + "_x"]
    for qw in (io.quant_weight, plan.quant_weight):
        wq, sigma = qw(w, 4, True)
        wd = wq / sigma.reshape(-1, 1, 1, 1)
        if tag == "pct_dw":
            xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
            y = sum(wd[:, 0, i, j].reshape(1, -1, 1, 1) * xp[:, :, i:i + 5, j:j + 5] for i in range(3) for j in range(3))
            y = y + g[tag + "_b"].reshape(1, -1, 1, 1)
        else:
            y = np.einsum("oc,bchw->bohw", wd.reshape(wd.shape[0], -1), x)
        np.testing.assert_allclose(y, g[tag + "_y"], rtol=1e-12, atol=1e-12)
    # the plain ranges give different integers for the wide rows: the option is not a no-op
    assert not np.array_equal(io.quant_weight(g["pct_wide_w"], 4, False)[0], io.quant_weight(g["pct_wide_w"], 4, True)[0])


def test_bn_fold_conv_kat(golden):
    g = golden("quant_kat.npz")
    w, b = io.fold_bn(g["bnconv_w"], g["bnconv_gamma"], g["bnconv_beta"], g["bnconv_mean"], g["bnconv_var"],
                      float(g["bnconv_eps"]))
    wq, sigma = io.quant_weight(w, 4)
    y = np.einsum("oc,bchw->bohw", (wq / sigma.reshape(-1, 1, 1, 1)).reshape(7, 5), g["bnconv_x"]) + b.reshape(1, -1, 1, 1)
    np.testing.assert_allclose(y, g["bnconv_y"], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("name", ["dw_s1", "dw_s2", "dense", "g2"])
def test_deform_op_kat(golden, name):
    g = golden("deform_kat.npz")
    st, pad, dil, groups, dg = g[name + "_cfg"]
    y = deform_ref.deform_conv(g[name + "_x"], g[name + "_off"], g[name + "_w"], st, pad, dil, groups, dg)
    np.testing.assert_allclose(y, g[name + "_y"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["mod_same", "mod_chan", "mod_s2"])
def test_codesigned_module_kat(golden, name):
    g = golden("deform_kat.npz")
    st, bound = g[name + "_cfg"]
    wc = g[name + "_wc"] if name + "_wc" in g.files else None
    y = deform_ref.codesigned_module(g[name + "_x"], g[name + "_ws"], g[name + "_bs"], g[name + "_w"], int(st), int(bound), wc)
    np.testing.assert_allclose(y, g[name + "_y"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["voc", "small"])
def test_decode_kat(golden, name):
    g = golden("decode_kat.npz")
    hm = g[name + "_hm"].astype(np.float64)
    logit = np.log(hm) - np.log1p(-hm)                   # any monotone map gives the same peaks and order
    K = int(g[name + "_K"])
    for reg, ref in ((g[name + "_reg"], g[name + "_dets"]), (None, g[name + "_dets_noreg"])):
        dets, inds = io.ctdet_decode(logit, g[name + "_wh"].astype(np.float64), None if reg is None else reg.astype(np.float64), K)
        dets[..., 4] = np.take_along_axis(hm.reshape(hm.shape[0], -1), inds, 1)
        np.testing.assert_allclose(dets, ref, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("mode", ["round", "bilinear"])
def test_int_oracle_reproduces_reference_fp64(golden, calib, mode):
    """Every int8 grid the reference produced in fp64 is reproduced bit for bit; head outputs to 1e-12."""
    g = golden("codenet1x_256_%s.npz" % mode)
    st = make_quant_state(CFG, calib, mode, 256)
    o = io.IntOracle(CFG, st, mode)
    out = o.forward(make_images(2, 256, seed=2))
    assert o.saturated == 0
    checked = 0
    for k in g.files:
        if g[k].dtype == np.int8:
            assert int8_mismatch(o.cap[k], g[k]) == 0, k
            checked += 1
    assert checked >= 10
    for i in range(3):
        np.testing.assert_allclose(o.cap["up%d.sval" % i], g["up%d.sval" % i], rtol=0, atol=1e-12)
    for n, k in (("hm", "hm_logit"), ("wh", "wh"), ("reg", "reg")):
        np.testing.assert_allclose(out[n], g[k], rtol=1e-12, atol=1e-12)
    check_reference_dets(g["dets"][:2], out["hm"], out["wh"], out["reg"])


DENSE_512 = ("stem", "layer1.out", "layer2.out", "layer3.out", "layer4", "up0.deform", "up0.out", "up1.deform", "up1.out",
             "up2.deform", "up2.out")


@pytest.mark.parametrize("mode", ["round", "bilinear"])
def test_int_oracle_512(golden, calib, mode):
    """Config c geometry, dense: every stage output, the whole deformable path, heads and detections of two images."""
    g = golden("codenet1x_512_%s.npz" % mode)
    st = make_quant_state(CFG, calib, mode, 512)
    o = io.IntOracle(CFG, st, mode)
    out = o.forward(make_images(2, 512, seed=3))
    assert o.saturated == 0
    for k in DENSE_512:
        assert int8_mismatch(o.cap[k], g[k]) == 0, k
    for n, k in (("hm", "hm_logit"), ("wh", "wh"), ("reg", "reg")):
        np.testing.assert_array_equal(out[n].astype(np.float32), g[k])
    check_reference_dets(g["dets"][:2], out["hm"], out["wh"], out["reg"])


def test_int_oracle_w2_maxpool_512(golden):
    """Config e geometry at 512^2 (first deformable layer C = 2153), dense."""
    from codenet_b200.arch import NetConfig
    from util import maxpool3s2_int8
    cfg = NetConfig(num_classes=20, w2=True, maxpool=True)
    g = golden("codenet_w2mp_512_round.npz")
    st = make_quant_state(cfg, golden("codenet_w2mp_calib.npz"), "round", 512)
    o = io.IntOracle(cfg, st, "round")
    out = o.forward(make_images(2, 512, seed=3)[:1])
    for k in DENSE_512:
        ref = maxpool3s2_int8(g[k]) if k == "stem" else g[k]
        assert int8_mismatch(o.cap[k], ref) == 0, k
    for n, k in (("hm", "hm_logit"), ("wh", "wh"), ("reg", "reg")):
        np.testing.assert_array_equal(out[n].astype(np.float32), g[k])
    check_reference_dets(g["dets"][:1], out["hm"], out["wh"], out["reg"])


def test_int_oracle_w2_maxpool(golden):
    """Config e geometry (2x width, stride-2 stem + MaxPool): grids, heads and detections of the fp64 reference."""
    from codenet_b200.arch import NetConfig
    from util import maxpool3s2_int8
    cfg = NetConfig(num_classes=20, w2=True, maxpool=True)
    g = golden("codenet_w2mp_256_round.npz")
    st = make_quant_state(cfg, golden("codenet_w2mp_calib.npz"), "round", 256)
    o = io.IntOracle(cfg, st, "round")
    out = o.forward(make_images(2, 256, seed=2)[:1])
    assert o.saturated == 0
    checked = 0
    for k in g.files:
        if g[k].dtype == np.int8:
            ref = maxpool3s2_int8(g[k]) if k == "stem" else g[k]
            assert int8_mismatch(o.cap[k], ref) == 0, k
            checked += 1
    assert checked >= 20
    for n, k in (("hm", "hm_logit"), ("wh", "wh"), ("reg", "reg")):
        np.testing.assert_allclose(out[n], g[k], rtol=1e-12, atol=1e-12)
    check_reference_dets(g["dets"][:1], out["hm"], out["wh"], out["reg"])


def test_tf32_split_properties():
    """oracle/tf32_split.py (the operand split of the tensor-core float GEMM): hi carries at most 11 significant bits, the fp32
    remainder is exact, |lo| <= 2^-11 |x|, the product of two hi parts is exactly representable in fp32, and x - hi - lo (what the
    GEMM drops) is below 2^-22 |x|."""
    from oracle import tf32_split as ts
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(0, 1, 20000), rng.normal(0, 1e-6, 1000), rng.normal(0, 1e6, 1000), [0.0, 1.0, -1.0, 1.0009765625,
                        np.float32(1 + 2 ** -11), np.float32(1 + 2 ** -11 + 2 ** -23)]]).astype(np.float32)
    hi, lo = ts.split(x)
    assert not (hi.view(np.uint32) & 0x1FFF).any() and not (lo.view(np.uint32) & 0x1FFF).any()
    rem = (x.astype(np.float64) - hi.astype(np.float64))
    assert (rem == (x - hi).astype(np.float64)).all()                                  # the fp32 subtraction is exact
    assert (np.abs(rem) <= np.abs(x.astype(np.float64)) * 2.0 ** -11).all()
    assert (np.abs(rem - lo) <= np.abs(x.astype(np.float64)) * 2.0 ** -22).all()
    w = rng.normal(0, 1, x.size).astype(np.float32)
    wh, _ = ts.split(w)
    p64 = hi.astype(np.float64) * wh.astype(np.float64)
    assert (p64 == (hi * wh).astype(np.float64)).all()                                 # hi * hi: 22-bit product, exact in fp32
    assert ts.tf32_rna(np.float32([1 + 2 ** -11]))[0] == np.float32(1 + 2 ** -10)      # ties away from zero
    assert ts.tiling(244) == (2, 122, 128) and ts.tiling(2153) == (17, 127, 128) and ts.tiling(2) == (1, 2, 16)
    assert ts.pack_weights(rng.normal(0, 1, (244, 61)).astype(np.float32)).size == 2 * 128 * 64 * 2
