"""Boundary B2 at module granularity: forward() of the compat.Quant* classes (portable_quantizer/quant_modules.py:202-225,
278-321, 364-419, 473-517, 668-671, 878-907, 1061-1071) on QTensor values, against the compiled engine, the reference-generated
vectors and the oracle."""
import numpy as np
import pytest

from codenet_b200.arch import NetConfig
from codenet_b200.synth import make_quant_state, make_images
from oracle import int_oracle as io
from util import int8_mismatch

pytestmark = pytest.mark.gpu
CFG = NetConfig(num_classes=20)


def _detector(calib, mode, res, cfg=CFG, **kw):
    from codenet_b200 import compat
    from codenet_b200.compat.detector import default_opt
    st = make_quant_state(cfg, calib, mode, res)
    opt = default_opt(state_dict=st, offset_mode=mode, input_h=res, input_w=res, w2=cfg.w2, maxpool=cfg.maxpool, **kw)
    return compat.CtdetDetector(opt), st


@pytest.mark.parametrize("mode", ["round", "bilinear"])
def test_module_by_module_forward_equals_engine_and_reference(golden, calib, mode):
    """The reference's forward (shufflenetv2_dcn.py:314-330) run module by module -- every Quant* forward launching its own
    kernels -- gives the int8 grids the reference recorded and the heads of the compiled engine, bit for bit."""
    import torch
    from codenet_b200.compat import QTensor, PendingConv
    g = golden("codenet1x_256_%s.npz" % mode)
    det, st = _detector(calib, mode, 256, max_batch=2)
    m = det.model
    x = torch.from_numpy(make_images(2, 256, seed=2)).cuda()
    m.offset_mode = mode
    # stage by stage, with the intermediate values checked against the reference's grids
    p = m.layer0[0](x)
    assert isinstance(p, PendingConv) and p.shape == (2, 24, 64, 64)
    a = m.layer0[1](p)
    assert isinstance(a, QTensor) and a.shape == (2, 24, 64, 64)
    assert int8_mismatch(a.int_values().cpu().numpy(), g["stem"]) == 0
    u = m.layer1[0](a)                                               # stride-2 unit
    assert u.shape == (2, 116, 32, 32) and u.half == 64
    for k, unit in (("layer1.out", m.layer1), ("layer2.out", m.layer2), ("layer3.out", m.layer3)):
        a = unit(a)
        if k in g.files:
            assert int8_mismatch(a.int_values().cpu().numpy(), g[k]) == 0, k
    a = m.layer4(a)
    assert int8_mismatch(a.int_values().cpu().numpy(), g["layer4"]) == 0
    for i in range(3):
        qd, act, up = m.deconv_layers[3 * i], m.deconv_layers[3 * i + 1], m.deconv_layers[3 * i + 2]
        qd.offset_mode = mode
        pend = qd(a)
        assert isinstance(pend, PendingConv)
        a = act(pend)
        assert int8_mismatch(a.int_values().cpu().numpy(), g["up%d.out" % i]) == 0, i
        a = up(a)
        assert a.up == 1
    heads = {h: getattr(m, h)(a) for h in m.heads}
    for h, key in (("hm", "hm_logit"), ("wh", "wh"), ("reg", "reg")):
        np.testing.assert_allclose(heads[h].cpu().numpy(), g[key], rtol=2e-7, atol=1e-7)
    # the whole thing in one call, against the compiled engine
    out_m = m.forward_modules(x)[-1]
    out_e = m(x)[-1]
    for h in m.heads:
        np.testing.assert_array_equal(out_m[h].cpu().numpy(), out_e[h].cpu().numpy(), err_msg=h)
    # the reference's fp32 view of an activation
    s, z = a.act
    np.testing.assert_allclose(a.dequantize().cpu().numpy(), (a.int_values().cpu().numpy().astype(np.float64) + z) / s, rtol=1e-6)


def test_quantact_on_real_values_and_leaf_modules(calib):
    """QuantAct.forward on fp32 values (quant_utils.py:31-39), QuantBnConv2d / Quant_Conv2d as leaves (their fp32 return value
    through .dequantize() / forward), and the loud failures."""
    import torch
    from codenet_b200 import compat
    from codenet_b200.plan import act_params
    det, st = _detector(calib, "round", 256, max_batch=1)
    m = det.model
    rng = np.random.default_rng(5)
    xr = rng.uniform(-3, 5, (2, 24, 16, 16)).astype(np.float32)
    qa = compat.QuantAct(8, quant_mode="asymmetric")
    qa.set_range(-2.5, 4.0)
    q = qa(torch.from_numpy(xr).cuda())
    s, z = act_params(-2.5, 4.0, 8)
    want = np.clip(np.rint(np.float64(s) * xr.astype(np.float64) - z), -128, 127)
    assert int8_mismatch(q.int_values().cpu().numpy(), want) == 0
    assert qa(q) is q                                                # already on this grid
    # a 1x1 QuantBnConv2d as a leaf: fp32 result of the conv itself = sum w_hat * x_hat + b' (quant_modules.py:364-419)
    a = m.layer0[1](m.layer0[0](torch.from_numpy(make_images(1, 256, seed=2)).cuda()))
    unit = m.layer1[0]
    pend = unit.quant_convbn1(a)
    y = pend.dequantize().cpu().numpy().astype(np.float64)
    o = io.IntOracle(CFG, st, "round")
    from codenet_b200.arch import build_graph
    c = build_graph(CFG).units[0]["convs"]["pw1"]
    wq, sigma, b = o.weights(c)
    xh = a.dequantize().cpu().numpy().astype(np.float64)
    ref = np.einsum("oc,bchw->bohw", (wq.reshape(c.cout, c.cin) / sigma.reshape(-1, 1)), xh) + b.reshape(1, -1, 1, 1)
    np.testing.assert_allclose(y, ref, rtol=1e-5, atol=1e-5)
    # closing it with ReLU + QuantAct gives the unit's first grid
    a1 = unit.quant_act1(torch.nn.functional.relu(pend))
    o.forward(make_images(1, 256, seed=2))
    assert int8_mismatch(a1.int_values().cpu().numpy(), o.cap["layer1.0.act1"]) == 0
    with pytest.raises(TypeError, match="takes a QTensor"):
        unit(torch.zeros(1, 24, 64, 64, device="cuda"))
    live = compat.QuantAct(8, quant_mode="asymmetric")
    with pytest.raises(RuntimeError, match="running statistic"):
        live(torch.zeros(1, 4, 4, 4, device="cuda"))


def test_quant_deform_conv2d_forward_matches_reference_vectors(golden):
    """QuantDeformConv2d.forward(x, offset) -- the general op behind quant_modules.py:473-517 -- against the reference's
    deform_conv on the same tensors (full precision), and with 4-bit per-channel weights against the op on dequantised ones."""
    import torch
    from codenet_b200 import compat
    from codenet_b200.plan import quant_weight
    g = golden("deform_kat.npz")
    x, off, w = g["dw_s1_x"], g["dw_s1_off"], g["dw_s1_w"]
    conv = compat.DeformConv(6, 6, 3, stride=1, padding=1, groups=6, bias=False)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(w))
    qd = compat.QuantDeformConv2d(4, quant_mode="symmetric", per_channel=True)
    qd.set_param(conv)
    qd = qd.cuda()
    xt, ot = torch.from_numpy(x).float().cuda(), torch.from_numpy(off).float().cuda()
    qd.full_precision_flag = True
    np.testing.assert_allclose(qd(xt, ot).cpu().numpy(), g["dw_s1_y"], rtol=1e-4, atol=1e-4)
    qd.full_precision_flag = False
    wq, sigma = quant_weight(w.astype(np.float64), 4)
    from oracle import deform_ref
    want = deform_ref.deform_conv(x, off, wq / sigma.reshape(-1, 1, 1, 1), 1, 1, 1, 6, 1)
    np.testing.assert_allclose(qd(xt, ot).cpu().numpy(), want, rtol=1e-4, atol=1e-4)


def test_wt_percentile_engine_equals_oracle(calib):
    """--wt-percentile (quant_modules.py:382-395) through the plan compiler and the engine: weight ranges from the 0.1 / 99.9
    percentiles (0.95 * min / max for the 3x3 depthwise kernels) give other integer weights; engine and oracle agree bit for bit."""
    import torch
    from codenet_b200.engine import Engine
    cfg = NetConfig(num_classes=20, wt_percentile=True)
    st = make_quant_state(cfg, {k: calib[k] for k in calib.files if k != "digest"}, "round", 256)
    eng = Engine.from_state_dict(cfg, st, 256, 256, 2, offset_mode="round")
    x = make_images(2, 256, seed=9)
    out = eng.run(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    o = io.IntOracle(cfg, st, "round")
    ref = o.forward(x)
    for lbl in ("stem", "layer2.out", "layer4", "up0.deform", "up2.out"):
        assert int8_mismatch(eng.read_logical(lbl, 2), o.cap[lbl]) == 0, lbl
    np.testing.assert_array_equal(eng.read_heads(2), np.concatenate([ref["hm"], ref["wh"], ref["reg"]], 1).astype(np.float32))
    plain = io.IntOracle(CFG, st, "round")
    plain.forward(x)
    assert int8_mismatch(plain.cap["layer4"], o.cap["layer4"]) > 0          # the option changes the network
    eng.close()
