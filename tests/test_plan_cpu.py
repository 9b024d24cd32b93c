"""Host logic without a GPU: the plan compiler against the oracle, and the C-ABI surface."""
import os
import re

import numpy as np
import pytest

import plan_sim
from codenet_b200 import _lib
from codenet_b200.arch import NetConfig, build_graph, raw_param_shapes, raw_to_quant_key, act_keys
from codenet_b200.plan import build_plan
from codenet_b200.synth import make_quant_state, make_images, make_raw_state, state_digest
from oracle import int_oracle as io
from util import int8_mismatch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = NetConfig(num_classes=20)


def test_synthetic_weights_match_calibration_archive(calib):
    assert state_digest(make_raw_state(CFG, 0)) == str(calib["digest"])


def test_key_spaces_are_consistent():
    g = build_graph(CFG)
    raw = raw_param_shapes(g)
    r2q = raw_to_quant_key(g)
    assert set(raw) == set(r2q) and len(set(r2q.values())) == len(r2q)
    assert len(act_keys(g)) == 1 + 16 * 2 + 3 + 3 + 1 + 9 + 6          # stem, units, act4 x3, shared x3, layer4, ups, heads
    # CoDeNet1x parameter count, SURVEY.md section 6: 1.588 M
    n = sum(int(np.prod(s)) for k, s in raw.items() if "running" not in k)
    assert abs(n - 1.588e6) < 0.01e6


@pytest.mark.parametrize("res", [256])
def test_plan_descriptors_reproduce_oracle(calib, res):
    """Interpreting the descriptors handed to the kernels (chunk tables = split/cat/shuffle, HALF layout, virtual
    upsampling, 1x1-before-upsample, fused heads) with integer arithmetic gives the oracle's grids exactly."""
    st = make_quant_state(CFG, calib, "round", res)
    x = make_images(2, res, seed=2)
    plan = build_plan(CFG, st, res, res, "round")
    o = io.IntOracle(CFG, st, "round")
    out = o.forward(x)
    T, heads = plan_sim.run_plan(plan, x)
    n = 0
    for lbl in plan.taps:
        if lbl in o.cap:
            assert int8_mismatch(plan_sim.logical(plan, T, lbl), o.cap[lbl]) == 0, lbl
            n += 1
    assert n >= 40
    np.testing.assert_array_equal(heads, np.concatenate([out["hm"], out["wh"], out["reg"]], 1))


def test_plan_descriptors_w2_maxpool(golden):
    """The same for the 2x-width, stride-2 + MaxPool configuration (odd interleave groups of 61 channels, two N tiles
    in stage 3, 2153-channel deformable layer)."""
    cfg = NetConfig(num_classes=20, w2=True, maxpool=True)
    st = make_quant_state(cfg, golden("codenet_w2mp_calib.npz"), "round", 256)
    x = make_images(2, 256, seed=2)[:1]
    plan = build_plan(cfg, st, 256, 256, "round")
    o = io.IntOracle(cfg, st, "round")
    out = o.forward(x)
    T, heads = plan_sim.run_plan(plan, x)
    n = 0
    for lbl in plan.taps:
        if lbl in o.cap:
            assert int8_mismatch(plan_sim.logical(plan, T, lbl), o.cap[lbl]) == 0, lbl
            n += 1
    assert n >= 40
    np.testing.assert_array_equal(heads, np.concatenate([out["hm"], out["wh"], out["reg"]], 1))


def test_plan_shapes_config_c(calib):
    st = make_quant_state(CFG, calib, "round", 512)
    plan = build_plan(CFG, st, 512, 512, "round")
    kinds = [op.kind for op in plan.ops]
    assert kinds.count("deform") == 3 and kinds.count("stem") == 1 and kinds.count("pw") == 41 and kinds.count("dw") == 20
    assert (plan.out_H, plan.out_W, plan.cat) == (128, 128, 20)
    # the three deformable layers of SURVEY.md F2: C=1024@16, 256@32, 128@64
    d = [(op.a["C"], plan.tensors[op.a["out_t"]].H) for op in plan.ops if op.kind == "deform"]
    assert d == [(1024, 16), (256, 32), (128, 64)]
    for t in plan.tensors:
        assert t.pitch % 32 == 0


def test_plan_rejects_bad_input():
    with pytest.raises(ValueError):
        build_plan(CFG, {}, 250, 256)
    with pytest.raises(ValueError):
        build_plan(CFG, {}, 256, 256, "nearest")


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "codenet_b200.h")).read()
    declared = set(re.findall(r"\b(cdn_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"cdn_stream_t"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "symbol %s declared in include/codenet_b200.h is not exported" % name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.cdn_version() >= 100


def test_workspace_and_packing_sizes_are_host_arithmetic():
    """The size queries of the C ABI need no device: ctdet decode workspace (8 bytes per heat-map element + one counter per image
    + 1), the deformable module's scale scratch (8 bytes per output pixel), and the packed weights of the tensor-core float GEMM
    (N tiles of <= 128 output channels, columns padded to 16, input channels padded to 16, hi + lo)."""
    lib = _lib.load()
    assert lib.cdn_ctdet_decode_ws_bytes(4, 20, 64, 64) == 4 * 20 * 64 * 64 * 8 + 5 * 4
    assert lib.cdn_ctdet_decode_ws_bytes(0, 20, 64, 64) == 8 + 4 and lib.cdn_ctdet_decode_ws_bytes(1, 0, 64, 64) == 0
    assert lib.cdn_deform_dw_f32_ws_bytes(3, 16, 16, 1) == 3 * 256 * 8 and lib.cdn_deform_dw_f32_ws_bytes(3, 16, 16, 2) == 3 * 64 * 8
    assert lib.cdn_deform_dw_f32_ws_bytes(3, 16, 16, 3) == 0
    for Co, Cin, nt, bn in ((122, 24, 1, 128), (244, 244, 2, 128), (2153, 976, 17, 128), (80, 64, 1, 80), (2, 64, 1, 16), (192, 64, 2, 96),
                            (300, 976, 3, 112)):
        kpad = (Cin + 15) // 16 * 16
        assert lib.cdn_pw_tf32x3_packed_floats(Co, Cin) == nt * bn * kpad * 2, (Co, Cin)
    assert lib.cdn_pw_tf32x3_packed_floats(0, 5) == 0


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.cdn_check_device(0) == -3
    assert b"no CPU fallback" in lib.cdn_last_error()
    import ctypes as C
    h = C.c_void_p()
    assert lib.cdn_engine_create(C.byref(h), 0) == -3


def test_plan_file_round_trip(calib, tmp_path):
    """The compiled-plan artefact (plan_io: one .npz, no pickle): every tensor record, op argument (arrays and fp64 scalars
    bit for bit) and tap survives save -> load."""
    from codenet_b200.plan import build_plan
    from codenet_b200.plan_io import save_plan, load_plan
    cfg = NetConfig(num_classes=20)
    st = make_quant_state(cfg, calib, "bilinear", 256)
    p = build_plan(cfg, st, 256, 256, "bilinear")
    f = str(tmp_path / "codenet1x_256.cdnplan.npz")
    save_plan(p, f)
    q = load_plan(f)
    assert q.cfg == p.cfg and (q.in_H, q.in_W, q.offset_mode, q.cat, q.out_H, q.out_W) == (p.in_H, p.in_W, p.offset_mode, p.cat, p.out_H, p.out_W)
    assert q.taps == p.taps and len(q.tensors) == len(p.tensors) and len(q.ops) == len(p.ops)
    for a, b in zip(p.tensors, q.tensors):
        assert (a.id, a.H, a.W, a.C, a.pitch, a.half, a.name) == (b.id, b.H, b.W, b.C, b.pitch, b.half, b.name)
        assert tuple(float(v) for v in a.act) == tuple(b.act)
    for a, b in zip(p.ops, q.ops):
        assert (a.kind, a.name) == (b.kind, b.name) and set(a.a) == set(b.a)
        for k, v in a.a.items():
            if isinstance(v, np.ndarray):
                assert v.dtype == b.a[k].dtype and np.array_equal(v, b.a[k]), (a.name, k)
            else:
                assert type(b.a[k]) in (int, float) and b.a[k] == v, (a.name, k)
    import os
    assert os.path.getsize(f) < 2.5e6                      # 1.6 M parameters as int8 + constants


def test_reference_written_checkpoint_loads(tmp_path):
    """A checkpoint written by the REFERENCE's own save_model (lib/models/model.py:91-100: {'epoch', 'state_dict'}) from its own
    quantised network loads into the compat detector's model through load_model (:35-88 semantics: 'module.' prefix stripped,
    optimizer entry ignored) with every tensor identical.  Needs the reference tree (build container)."""
    from oracle import ref_harness as H
    if not H.available():
        pytest.skip("reference tree not available")
    import torch
    from codenet_b200 import compat
    from codenet_b200.synth import make_raw_state
    H.load_reference()
    from models.model import save_model
    cfg = NetConfig(num_classes=20)
    raw = make_raw_state(cfg, 0)
    ref = H.build_reference_model({k: torch.from_numpy(v) for k, v in raw.items()}, dict(cfg.head_list()), dtype=torch.float32)
    ref.eval()
    H.quantize_reference_model(ref)
    with torch.no_grad():
        for n, b in ref.named_buffers():                   # give the QuantAct ranges recognisable values
            if n.endswith("x_min"):
                b.fill_(-1.25)
            if n.endswith("x_max"):
                b.fill_(3.5)
    f = str(tmp_path / "model_last.pth")
    save_model(f, 7, torch.nn.DataParallel(ref) if False else ref)
    ck = torch.load(f, map_location="cpu")
    assert set(ck) >= {"epoch", "state_dict"}
    # as a DataParallel training run would have written it
    torch.save({"epoch": 7, "state_dict": {"module." + k: v for k, v in ck["state_dict"].items()}, "optimizer": {}}, f)
    det = compat.CtdetDetector.__new__(compat.CtdetDetector)
    det.model = compat.PoseShuffleNetV2({"hm": 20, "wh": 2, "reg": 2}, 64)
    compat.quantize_shufflenetv2_dcn(det.model, 4, None, 8, "symmetric", "asymmetric", True, False, False, False)
    unexpected = det.load_model(f)
    assert not unexpected
    mine, theirs = det.model.state_dict(), ref.state_dict()
    assert set(mine) == set(theirs)
    for k in theirs:
        assert torch.equal(mine[k].float(), theirs[k].float()), k
