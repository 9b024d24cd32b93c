"""GPU parity of the reference-facing interface (codenet_b200.compat): the calls a user of the reference makes."""
import numpy as np
import pytest

from codenet_b200.arch import NetConfig
from codenet_b200.synth import make_quant_state, make_images
from oracle import int_oracle as io
from util import assert_deform_f32_close

pytestmark = pytest.mark.gpu
CFG = NetConfig(num_classes=20)


def test_deform_conv_function_matches_reference_vectors(golden):
    """`deform_conv(x, offset, weight, stride, padding, dilation, groups, deformable_groups)` -- boundary B1."""
    import torch
    from codenet_b200 import compat
    g = golden("deform_kat.npz")
    for name in ("dw_s1", "dw_s2", "dense", "g2"):
        x, off, w, y = (torch.from_numpy(g[name + s].astype(np.float32)).cuda() for s in ("_x", "_off", "_w", "_y"))
        st, pad, dil, groups, dg = (int(v) for v in g[name + "_cfg"])
        out = compat.deform_conv(x, off, w, st, pad, dil, groups, dg)
        assert out.shape == y.shape
        # every element within 1e-4 relative; floor()-boundary flips (if any) explicitly bounded, none ignored
        assert_deform_f32_close(out.cpu().numpy(), g[name + "_x"], g[name + "_off"], g[name + "_w"], g[name + "_y"], st, pad, dil,
                                groups, dg, name)
    with pytest.raises(AssertionError, match="im2col step must divide batchsize"):
        compat.deform_conv(torch.zeros(3, 4, 5, 5).cuda(), torch.zeros(3, 18, 5, 5).cuda(), torch.zeros(4, 1, 3, 3).cuda(),
                           1, 1, 1, 4, 1, 2)


@pytest.mark.parametrize("name", ["mod_same", "mod_chan", "mod_s2"])
def test_codesigned_module_fp32(golden, name):
    """DeformConvWithOffsetScaleBoundPositive.forward (fp32, bilinear offsets) within 1e-4 relative of the reference."""
    import torch
    from codenet_b200 import compat
    g = golden("deform_kat.npz")
    st, bound = (int(v) for v in g[name + "_cfg"])
    x = g[name + "_x"]
    cin, cout = x.shape[1], g[name + "_y"].shape[1]
    m = compat.DeformConvWithOffsetScaleBoundPositive(cin, cout, 3, st, 1, groups=cout, offset_bound=bound).cuda()
    with torch.no_grad():
        m.conv_scale.weight.copy_(torch.from_numpy(g[name + "_ws"])); m.conv_scale.bias.copy_(torch.from_numpy(g[name + "_bs"]))
        m.conv.weight.copy_(torch.from_numpy(g[name + "_w"]))
        if cin != cout:
            m.conv_channel.weight.copy_(torch.from_numpy(g[name + "_wc"]))
        y = m(torch.from_numpy(x.astype(np.float32)).cuda())          # cdn_deform_dw_f32 (+ cdn_pw_f32): no library conv
    # against the fp64 oracle evaluated on the fp32-rounded parameters: the module's offsets are anchor * (s - 1) with s a
    # continuous function of x, so floor() flips need |frac| below fp32 resolution; every element is bounded
    from oracle import deform_ref
    f64 = np.float64
    r32 = lambda a: np.asarray(a, np.float32).astype(f64)
    ref = deform_ref.codesigned_module(r32(x), r32(g[name + "_ws"]), r32(g[name + "_bs"]), r32(g[name + "_w"]), st, bound,
                                       r32(g[name + "_wc"]) if cin != cout else None)
    np.testing.assert_allclose(ref, g[name + "_y"], rtol=0, atol=2e-5 * max(1.0, np.abs(g[name + "_y"]).max()))
    err = np.abs(y.cpu().numpy().astype(f64) - ref)
    scale = max(1.0, float(np.abs(ref).max()))
    s_map = np.tensordot(r32(x)[:, :, ::st, ::st], r32(g[name + "_ws"]).reshape(-1), axes=(1, 0)) + float(g[name + "_bs"].reshape(-1)[0])
    s_map = np.clip(s_map, -bound + 1, bound)
    near = (np.abs(s_map - np.rint(s_map)) < 2e-5)[:, None]            # taps land on integer rows / columns only when s is integral
    bad = err > 1e-4 * scale
    assert not (bad & ~near).any(), (name, int((bad & ~near).sum()), float(err.max()))
    if bad.any():
        cap = 4e-5 * float(np.abs(g[name + "_w"]).reshape(cin, -1).sum(1).max()) * float(np.abs(x).max()) * \
              (float(np.abs(g[name + "_wc"]).sum(1).max()) if cin != cout else 1.0) + 1e-4 * scale
        assert err[bad].max() <= cap, (name, float(err[bad].max()), cap)


@pytest.mark.parametrize("name", ["voc", "small"])
def test_ctdet_decode_on_probabilities(golden, name):
    """ctdet_decode(heat, wh, reg, K) on the reference's own post-sigmoid vectors (lib/models/decode.py:474-505)."""
    import torch
    from codenet_b200 import compat
    g = golden("decode_kat.npz")
    K = int(g[name + "_K"])
    hm, wh, reg = (torch.from_numpy(g[name + s]).cuda() for s in ("_hm", "_wh", "_reg"))
    d = compat.ctdet_decode(hm, wh, reg=reg, K=K).cpu().numpy()
    ref = g[name + "_dets"]
    np.testing.assert_array_equal(d[..., 5], ref[..., 5])                 # classes in the reference's order (tie-free vectors)
    np.testing.assert_array_equal(d[..., 4], ref[..., 4])                 # scores are the given probabilities, untouched
    np.testing.assert_allclose(d[..., :4], ref[..., :4], rtol=1e-6, atol=1e-5)
    d2 = compat.ctdet_decode(hm, wh, reg=None, K=K).cpu().numpy()
    np.testing.assert_allclose(d2[..., :4], g[name + "_dets_noreg"][..., :4], rtol=1e-6, atol=1e-5)


def _detector(calib, mode, res, **kw):
    from codenet_b200 import compat
    from codenet_b200.compat.detector import default_opt
    st = make_quant_state(CFG, calib, mode, res)
    opt = default_opt(state_dict=st, offset_mode=mode, input_h=res, input_w=res, **kw)
    return compat.CtdetDetector(opt)


@pytest.mark.parametrize("mode", ["round", "bilinear"])
def test_detector_process_matches_reference_vectors(golden, calib, mode):
    """CtdetDetector.process(images) -- boundary B3 -- against the fp64 run of the unmodified reference."""
    import torch
    g = golden("codenet1x_256_%s.npz" % mode)
    det = _detector(calib, mode, 256, max_batch=2)
    x = torch.from_numpy(make_images(2, 256, seed=2)).cuda()
    output, dets = det.process(x)
    np.testing.assert_allclose(output["hm"].cpu().numpy(), 1 / (1 + np.exp(-g["hm_logit"])), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(output["wh"].cpu().numpy(), g["wh"], rtol=2e-7, atol=1e-7)
    np.testing.assert_allclose(output["reg"].cpu().numpy(), g["reg"], rtol=2e-7, atol=1e-7)
    h64 = np.concatenate([g["hm_logit"], g["wh"], g["reg"]], 1).astype(np.float32).astype(np.float64)
    odets, _ = io.ctdet_decode(h64[:, :20], h64[:, 20:22], h64[:, 22:24], 100)
    np.testing.assert_allclose(dets.cpu().numpy(), odets, rtol=1e-5, atol=1e-4)
    # the network mirror returns logits like PoseShuffleNetV2.forward
    out = det.model(x)[-1]
    np.testing.assert_allclose(out["hm"].cpu().numpy(), g["hm_logit"], rtol=2e-7, atol=1e-7)


def test_detector_run_and_flip_test(calib):
    """run() on a raw image: same result keys as the reference (base_detector.py:153-155); flip_test averages the
    mirrored pass (ctdet.py:35-38) and must agree with doing that by hand from two plain passes."""
    import torch
    pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    det = _detector(calib, "round", 256, max_batch=2)
    ret = det.run(img)
    assert set(ret) == {"results", "tot", "load", "pre", "net", "dec", "post", "merge"}
    assert sorted(ret["results"]) == list(range(1, 21)) and sum(len(v) for v in ret["results"].values()) == 100
    assert all(v.shape[1] == 5 and v.dtype == np.float32 for v in ret["results"].values())
    det.opt.flip_test = True
    images, meta = det.pre_process(img, 1.0)
    assert images.shape == (2, 3, 256, 256)
    output, dets = det.process(images.cuda())
    det.opt.flip_test = False
    o2, _ = det.process(images.cuda())
    hm = (o2["hm"][0:1] + torch.flip(o2["hm"][1:2], [3])) / 2
    wh = (o2["wh"][0:1] + torch.flip(o2["wh"][1:2], [3])) / 2
    from codenet_b200 import compat
    want = compat.ctdet_decode(hm, wh, reg=o2["reg"][0:1], K=100)
    assert dets.shape == (1, 100, 6)
    np.testing.assert_array_equal(dets.cpu().numpy(), want.cpu().numpy())


def test_uint8_input_is_bit_identical_to_normalised_fp32(calib):
    """cdn_engine_run_u8 / run_host_u8 (normalisation table in the stem kernel) against the fp32 path on the tensor the
    reference's pre_process would build; and CtdetDetector.run() taking the fast path on an image at the input size."""
    import torch
    from codenet_b200.engine import Engine
    st = make_quant_state(CFG, calib, "round", 256)
    eng = Engine.from_state_dict(CFG, st, 256, 256, 3, offset_mode="round")
    mean, std = np.array([0.485, 0.456, 0.406], np.float32), np.array([0.229, 0.224, 0.225], np.float32)
    eng.set_normalization(mean, std)
    u8 = np.random.default_rng(7).integers(0, 256, (3, 256, 256, 3), dtype=np.uint8)
    f32 = ((u8 / 255. - mean.reshape(1, 1, 1, 3)) / std.reshape(1, 1, 1, 3)).astype(np.float32).transpose(0, 3, 1, 2).copy()
    a = eng.run(torch.from_numpy(f32).cuda()); a = {k: v.clone() for k, v in a.items()}
    ga = eng.read_logical("stem", 3)
    b = eng.run(torch.from_numpy(u8).cuda())
    np.testing.assert_array_equal(eng.read_logical("stem", 3), ga)
    for k in ("hm", "wh", "reg", "dets", "inds"):
        np.testing.assert_array_equal(a[k].cpu().numpy(), b[k].cpu().numpy(), err_msg=k)
    d_host, i_host = eng.run_host(u8)
    np.testing.assert_array_equal(d_host, a["dets"].cpu().numpy())
    np.testing.assert_array_equal(i_host, a["inds"].cpu().numpy())
    eng.close()
    pytest.importorskip("cv2")
    det = _detector(calib, "round", 256, max_batch=1)
    img = u8[0]
    fast = det.run(img)
    images, meta = det.pre_process(img, 1.0)
    _, dets = det.process(images.cuda())
    slow = det.merge_outputs([det.post_process(dets, meta, 1.0)])
    for j in range(1, 21):
        np.testing.assert_array_equal(fast["results"][j], slow[j])


def test_post_process_on_device_matches_host_restatement(calib):
    """cdn_ctdet_post_affine against the numpy restatement of ctdet_post_process (lib/utils/post_process.py:86-103)."""
    import torch
    from codenet_b200.compat import detector as D
    det = _detector(calib, "round", 256, max_batch=1)
    rng = np.random.default_rng(4)
    dets = np.zeros((1, 100, 6), np.float32)
    dets[0, :, :4] = rng.uniform(-5, 70, (100, 4)); dets[0, :, 4] = np.sort(rng.uniform(0, 1, 100))[::-1]
    dets[0, :, 5] = rng.integers(0, 20, 100)
    meta = {'c': np.array([213.5, 160.0], np.float32), 's': 427.0, 'out_height': 64, 'out_width': 64}
    got = det.post_process(torch.from_numpy(dets).cuda(), meta, 1.0)
    want = D.ctdet_post_process(dets.copy(), [meta['c']], [meta['s']], 64, 64, 20)[0]
    for j in range(1, 21):
        w = np.array(want[j], dtype=np.float32).reshape(-1, 5)
        np.testing.assert_array_equal(got[j], w)


def test_device_pre_process_equals_opencv_path(calib):
    """cdn_warp_affine_u8 against the oracle (pinned to cv2.warpAffine by tests/test_prepost_cpu.py) on
    landscape, portrait and input-sized frames with and without the --flip_test mirror; and run() through the device
    pre-process against run() through the reference's host pre_process (cv2), detection for detection."""
    import torch
    from oracle import warp_ref
    from codenet_b200.compat.detector import get_affine_transform
    det = _detector(calib, "round", 256, max_batch=2)
    rng = np.random.default_rng(8)
    for (h, w), flip in (((300, 400), False), ((400, 300), True), ((256, 256), False), ((97, 203), True)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        det.opt.flip_test = flip
        got, meta = det.pre_process_device(img, 1)
        c, s = np.array([w / 2., h / 2.], np.float32), max(h, w) * 1.0
        want = warp_ref.warp_affine(img, get_affine_transform(c, s, 0, [256, 256]), (256, 256))
        g = got.cpu().numpy()
        np.testing.assert_array_equal(g[0], want)
        if flip:
            np.testing.assert_array_equal(g[1], want[:, ::-1])
        assert meta["out_height"] == 64 and float(meta["s"]) == s
    pytest.importorskip("cv2")
    img = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    for flip in (False, True):
        det.opt.flip_test = flip
        det.opt.device_pre_process = True
        a = det.run(img)["results"]
        det.opt.device_pre_process = False
        b = det.run(img)["results"]
        for j in range(1, 21):
            np.testing.assert_array_equal(a[j], b[j])
    det.opt.flip_test = False


def test_group_by_class_on_device():
    """cdn_ctdet_group_by_class against a stable numpy grouping (lib/utils/post_process.py:86-103), out-of-range classes last."""
    import ctypes as C
    import torch
    from codenet_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(6)
    B, K, NC = 3, 100, 20
    dets = rng.uniform(0, 100, (B, K, 6)).astype(np.float32)
    dets[..., 5] = rng.integers(0, NC, (B, K))
    dets[1, 7, 5] = 25.0; dets[2, 3, 5] = -1.0
    d = torch.from_numpy(dets).cuda(); out = torch.zeros_like(d); cnt = torch.zeros((B, NC), dtype=torch.int32, device="cuda")
    _lib.check(L.cdn_ctdet_group_by_class(C.c_void_p(d.data_ptr()), B, K, NC, C.c_void_p(out.data_ptr()), C.c_void_p(cnt.data_ptr()),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    o, c = out.cpu().numpy(), cnt.cpu().numpy()
    for b in range(B):
        cls = dets[b, :, 5]
        key = np.where((cls >= 0) & (cls < NC), cls, NC)
        order = np.argsort(key, kind="stable")
        np.testing.assert_array_equal(o[b], dets[b][order])
        np.testing.assert_array_equal(c[b], np.bincount(key.astype(int), minlength=NC + 1)[:NC])


def test_ctdet_decode_edge_cases_from_the_reference(golden):
    """The reference's peak test runs on fp32 PROBABILITIES (decode.py:10-16 after hm.sigmoid_()): vectors generated from it
    with (i) fewer than K positive peaks -- the reference fills the K rows with score-0 entries in topk's arbitrary order, only
    the positive rows are defined -- and (ii) saturated neighbours whose distinct logits all map to probability 1.0f, which
    `hmax == heat` keeps as separate peaks (rows inside the tie group in unspecified order)."""
    import torch
    from codenet_b200 import compat
    g = golden("decode_kat.npz")
    hm, wh, reg = (torch.from_numpy(g["sparse" + s]).cuda() for s in ("_hm", "_wh", "_reg"))
    d = compat.ctdet_decode(hm, wh, reg=reg, K=int(g["sparse_K"])).cpu().numpy()[0]
    ref = g["sparse_dets"][0]
    n = int((ref[:, 4] > 0).sum())
    assert n == 25 and (d[:n, 4] > 0).all() and (d[n:, 4] == 0).all() and (ref[n:, 4] == 0).all()
    np.testing.assert_array_equal(d[:n, 4:], ref[:n, 4:])
    np.testing.assert_allclose(d[:n, :4], ref[:n, :4], rtol=1e-6, atol=1e-5)
    hm, wh, reg = (torch.from_numpy(g["saturated" + s]).cuda() for s in ("_hm", "_wh", "_reg"))
    d = compat.ctdet_decode(hm, wh, reg=reg, K=int(g["saturated_K"])).cpu().numpy()[0]
    ref = g["saturated_dets"][0]
    np.testing.assert_array_equal(d[:, 4], ref[:, 4])                     # the score list, ties included
    assert (ref[:4, 4] == 1.0).all()                                      # three saturated neighbours + one isolated peak
    key = lambda r: tuple(np.round(r.astype(np.float64), 3))
    i = 0
    while i < len(ref):                                                   # rows as multisets inside every group of equal score
        j = i
        while j + 1 < len(ref) and ref[j + 1, 4] == ref[i, 4]:
            j += 1
        assert sorted(map(key, d[i:j + 1])) == sorted(map(key, ref[i:j + 1])), (i, j)
        i = j + 1


@pytest.mark.parametrize("gemm", ["fp32", "tf32x3"])
def test_detector_float_model(golden, gemm):
    """CtdetDetector with resume_quantize=False (the reference's float test.py run) routes through EngineF32: the heads equal the
    fp64 reference run within the float-path bounds of tests/test_gpu_f32.py, `hm` comes back post-sigmoid (ctdet.py:32), and
    flip_test averages the mirrored pass like ctdet.py:35-38."""
    import torch
    from codenet_b200 import compat
    from codenet_b200.compat.detector import default_opt
    from codenet_b200.synth import make_raw_state
    g = golden("codenet_float_1x_256.npz")
    raw = make_raw_state(CFG, 0)
    for k in g.files:
        if k.startswith("bn/"):
            raw[k[3:]] = g[k]
    opt = default_opt(resume_quantize=False, state_dict={k: torch.from_numpy(np.asarray(v)) for k, v in raw.items()}, input_h=256, input_w=256,
                      f32_gemm=gemm)
    det = compat.CtdetDetector(opt)
    x = torch.from_numpy(make_images(2, 256, seed=2)).cuda()
    output, dets = det.process(x[:1])
    slack = 1.5 if gemm == "tf32x3" else 1.0
    for name, ref in (("hm", 1 / (1 + np.exp(-g["hm_logit"].astype(np.float64)))), ("wh", g["wh"].astype(np.float64)), ("reg", g["reg"].astype(np.float64))):
        got = output[name].cpu().numpy().astype(np.float64)
        l2 = np.sqrt(((got - ref) ** 2).sum() / (ref ** 2).sum())
        assert l2 <= max(1e-4, slack * float(g["ref_fp32_l2rel/" + name])), (name, l2)
    assert dets.shape == (1, 100, 6)
    np.testing.assert_allclose(np.sort(dets[0, :, 4].cpu().numpy())[::-1][:50], np.sort(g["dets"][0][:, 4])[::-1][:50], rtol=1e-3)
    det.opt.flip_test = True
    pair = torch.cat([x[:1], torch.flip(x[:1], [3])])
    o2, d2 = det.process(pair)
    det.opt.flip_test = False
    o3, _ = det.process(pair)
    want_hm = (o3["hm"][0:1] + torch.flip(o3["hm"][1:2], [3])) / 2
    assert torch.equal(o2["hm"], want_hm) and d2.shape == (1, 100, 6)
