"""Host-side mirror of the reference interface (codenet_b200.compat) without a GPU: key spaces against the reference's
own state dicts, checkpoint ingestion into the plan compiler, error behaviour, pre/post-processing arithmetic."""
import json
import os

import numpy as np
import pytest
import torch

from codenet_b200 import compat
from codenet_b200.arch import NetConfig
from codenet_b200.plan import build_plan
from codenet_b200.synth import make_quant_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADS = {"hm": 20, "wh": 2, "reg": 2}


def _quantised(w2=False, maxpool=False):
    m = compat.PoseShuffleNetV2(HEADS, 64, w2=w2, maxpool=maxpool)
    raw = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    compat.quantize_shufflenetv2_dcn(m, 4, None, 8, 'symmetric', 'asymmetric', True, False, False, False, w2=w2,
                                     maxpool=maxpool)
    return m, raw


@pytest.mark.parametrize("tag,w2,mp", [("1x", False, False), ("w2", True, False), ("1x_maxpool", False, True)])
def test_state_dict_keys_equal_the_reference(tag, w2, mp):
    """tests/golden/ref_state_keys.json was dumped from the UNMODIFIED reference (oracle/make_golden.py ref_keys)."""
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_state_keys.json")))
    m, raw = _quantised(w2, mp)
    assert raw == {k: tuple(v) for k, v in ref[tag + "/raw"].items()}
    q = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert q == {k: tuple(v) for k, v in ref[tag + "/quant"].items()}


def test_checkpoint_roundtrip_gives_the_same_plan(calib):
    """A state dict in the reference's quantised key space loads into the mirror and compiles to the same plan."""
    cfg = NetConfig(num_classes=20)
    st = make_quant_state(cfg, calib, "round", 256)
    m, _ = _quantised()
    det_like = compat.CtdetDetector.__new__(compat.CtdetDetector)
    det_like.model = m
    unexpected = compat.CtdetDetector.load_state_dict(det_like, st)
    assert not unexpected
    compat.freeze_ranges(m)
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    a, b = build_plan(cfg, st, 256, 256, "round"), build_plan(cfg, sd, 256, 256, "round")
    assert len(a.ops) == len(b.ops)
    for oa, ob in zip(a.ops, b.ops):
        assert oa.kind == ob.kind and oa.name == ob.name
        for k, v in oa.a.items():
            if isinstance(v, np.ndarray):
                np.testing.assert_array_equal(v, ob.a[k], err_msg="%s.%s" % (oa.name, k))
            else:
                assert v == ob.a[k], (oa.name, k)


def test_modules_refuse_eager_and_cpu_execution():
    m, _ = _quantised()
    x = torch.zeros(1, 3, 64, 64)
    with pytest.raises(RuntimeError, match="no CPU execution path"):
        m(x)
    with pytest.raises(TypeError, match="takes a QTensor"):            # module-level forwards run on QTensor values ...
        m.layer1[0](torch.zeros(1, 24, 16, 16))
    with pytest.raises(RuntimeError, match="running statistic"):       # ... need frozen ranges ...
        compat.QuantAct(8, quant_mode="asymmetric")(x)
    qa = compat.QuantAct(8, quant_mode="asymmetric"); qa.set_range(-1.0, 1.0)
    with pytest.raises(RuntimeError, match="no CPU execution path"):   # ... and a GPU
        qa(x)
    with pytest.raises(compat.NotCompiledError):
        compat.QuantLinear(4, 8, 8)(torch.zeros(1, 8))
    with pytest.raises(NotImplementedError):                 # the reference refuses CPU tensors the same way
        compat.deform_conv(torch.zeros(1, 4, 5, 5), torch.zeros(1, 18, 5, 5), torch.zeros(4, 1, 3, 3), 1, 1, 1, 4, 1)
    with pytest.raises(ValueError, match="Expected 4D tensor"):
        compat.deform_conv(torch.zeros(4, 5, 5), torch.zeros(1, 18, 5, 5), torch.zeros(4, 1, 3, 3))
    with pytest.raises(RuntimeError, match="no CPU execution path"):
        compat.ctdet_decode(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(ValueError, match="unknown quant mode"):
        compat.QuantAct(8, quant_mode="bogus")


def test_running_ranges_are_rejected():
    m, _ = _quantised()
    with pytest.raises(RuntimeError, match="running statistics"):
        m.compile_engine(64, 64, 1)


def test_reference_api_quirks():
    with pytest.raises(AssertionError):
        compat.DeformConv(4, 4, bias=True)                                   # modules/dcn_deform_conv.py:26
    d = compat.DeformConvWithOffsetScaleBoundPositive(8, 4, 3, 1, 1, groups=4, offset_bound=3)
    assert d.conv.groups == 8 and d.conv.out_channels == 8                   # `groups` ignored: always depthwise in->in
    assert float(d.conv_scale.weight.abs().sum()) == 0.0 and float(d.conv_scale.bias) == 1.0
    assert (d.conv_bound.min_val, d.conv_bound.max_val) == (-2, 3)
    assert tuple(d.anchor_offset.shape) == (1, 18, 1, 1) and not any(k.startswith("anchor") for k in d.state_dict())
    assert tuple(compat.QuantAct(8).x_min.shape) == (1,)
    y = compat.channel_shuffle(torch.arange(8.).view(1, 8, 1, 1), 2).flatten().tolist()
    assert y == [0, 4, 1, 5, 2, 6, 3, 7]                                     # out[2k] = x1[k], out[2k+1] = x2[k]


def test_affine_helpers_match_opencv():
    cv2 = pytest.importorskip("cv2")
    from codenet_b200.compat import detector as D
    c = np.array([320., 213.5], np.float32)
    for s, out in [(640.0, (128, 128)), (np.array([512., 384.], np.float32), (128, 96))]:
        for inv in (0, 1):
            got = D.get_affine_transform(c, s, 0, out, inv=inv)
            sc = np.array([s, s], np.float32) if not isinstance(s, np.ndarray) else s
            src = np.zeros((3, 2), np.float32); dst = np.zeros((3, 2), np.float32)
            src[0] = c; src[1] = c + np.array([0, sc[0] * -0.5], np.float32)
            dst[0] = [out[0] * 0.5, out[1] * 0.5]; dst[1] = dst[0] + np.array([0, out[0] * -0.5], np.float32)
            src[2] = D._third_point(src[0], src[1]); dst[2] = D._third_point(dst[0], dst[1])
            want = cv2.getAffineTransform(dst, src) if inv else cv2.getAffineTransform(src, dst)
            np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)


def test_post_process_groups_by_class():
    from codenet_b200.compat import detector as D
    rng = np.random.default_rng(0)
    dets = np.zeros((1, 10, 6), np.float32)
    dets[0, :, :4] = rng.uniform(0, 128, (10, 4)); dets[0, :, 4] = np.linspace(0.9, 0.1, 10)
    dets[0, :, 5] = rng.integers(0, 3, 10)
    c, s = np.array([256., 256.], np.float32), 512.0
    out = D.ctdet_post_process(dets.copy(), [c], [s], 128, 128, 3)[0]
    assert sorted(out) == [1, 2, 3] and sum(len(v) for v in out.values()) == 10
    # 128 output pixels span 512 image pixels: a pure x4 scaling
    k = int(dets[0, 0, 5]) + 1
    np.testing.assert_allclose(np.array(out[k][0][:4]), dets[0, 0, :4] * 4, rtol=1e-5)


def test_shard_bounds():
    from codenet_b200.shard import shard_bounds, my_slice
    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]
    assert my_slice(2048, 7, 8) == (1792, 2048)
    with pytest.raises(ValueError):
        my_slice(8, 8, 8)


def test_pre_process_is_normalisation_only_at_input_size():
    """For an image already at the input size the reference's resize + warpAffine (base_detector.py:61-65) are the
    identity, so the uint8 fast path (normalisation table inside the stem kernel) sees exactly pre_process's tensor."""
    pytest.importorskip("cv2")
    from codenet_b200.compat.detector import default_opt
    det = compat.CtdetDetector.__new__(compat.CtdetDetector)
    det.opt = default_opt(input_h=96, input_w=96)
    det.mean = np.array(det.opt.mean, np.float32).reshape(1, 1, 3)
    det.std = np.array(det.opt.std, np.float32).reshape(1, 1, 3)
    img = np.random.default_rng(1).integers(0, 256, (96, 96, 3), dtype=np.uint8)
    images, meta = compat.CtdetDetector.pre_process(det, img, 1.0)
    # the table cdn_engine_set_normalization builds: fl32((u/255. - mean)/std) evaluated in double
    lut = ((np.arange(256, dtype=np.float64)[None, :] / 255.0 - det.mean.reshape(3, 1).astype(np.float64))
           / det.std.reshape(3, 1).astype(np.float64)).astype(np.float32)
    want = np.stack([lut[c][img[:, :, c]] for c in range(3)])[None]
    np.testing.assert_array_equal(images.numpy(), want)
    assert meta["out_height"] == 24 and float(meta["s"]) == 96.0


def test_post_process_matches_the_reference(golden):
    """The numpy restatement of ctdet_post_process against vectors from the reference's own function (cv2 affine)."""
    from codenet_b200.compat import detector as D
    g = golden("post_kat.npz")
    out = D.ctdet_post_process(g["dets"].copy(), [g["c0"], g["c1"]], [float(g["s0"]), g["s1"]], 128, 128, 5)
    for i in range(2):
        for j in range(1, 6):
            got = np.array(out[i][j], np.float32).reshape(-1, 5)
            np.testing.assert_allclose(got, g["img%d_cls%d" % (i, j)], rtol=1e-6, atol=1e-4)


def test_soft_nms_matches_the_compiled_reference(golden):
    """compat.soft_nms against the reference's own Cython soft_nms (lib/models/external/nms.pyx:77-170), fixtures from
    oracle/make_golden_nms.py: the in-place result (reordered boxes, decayed scores, discarded tail) and the kept count
    must be bit-identical for all three methods."""
    from codenet_b200.compat import soft_nms
    g = golden("soft_nms_kat.npz")
    n = 0
    for key in g.files:
        if not key.endswith("_in"):
            continue
        tag = key[:-3]
        method = int(tag.split("_m")[1])
        a = g[key].copy()
        keep = soft_nms(a, Nt=0.5, method=method)
        np.testing.assert_array_equal(a, g[tag + "_out"], err_msg=tag)
        assert len(keep) == int(g[tag + "_keep"]), tag
        n += 1
    assert n == 36


def test_merge_outputs_multi_scale_uses_soft_nms():
    """Two scales of the same detections: merge_outputs (ctdet.py:59-74) decays the duplicates instead of raising."""
    from codenet_b200.compat.detector import CtdetDetector
    det = CtdetDetector.__new__(CtdetDetector)
    det.num_classes, det.max_per_image, det.scales = 2, 100, [1.0, 1.5]
    det.opt = type("O", (), {"nms": False})()
    a = {1: np.array([[10, 10, 50, 50, 0.9], [100, 100, 140, 150, 0.8]], np.float32), 2: np.zeros((0, 5), np.float32)}
    b = {1: np.array([[11, 10, 51, 50, 0.7]], np.float32), 2: np.array([[5, 5, 20, 20, 0.3]], np.float32)}
    r = det.merge_outputs([a, b])
    assert r[1].shape == (3, 5) and r[2].shape == (1, 5)
    assert r[1][0, 4] == np.float32(0.9) and r[1][1, 4] == np.float32(0.8)         # sorted by score, untouched
    assert 0 < r[1][2, 4] < np.float32(0.7) * np.float32(0.2)                      # near-duplicate decayed by exp(-iou^2/0.5)
