"""Whole-network parity on the GPU: engine vs the vectors of the unmodified reference (fp64) and vs the oracle."""
import numpy as np
import pytest

from codenet_b200 import _lib
from codenet_b200.arch import NetConfig
from codenet_b200.engine import Engine
from codenet_b200.synth import make_quant_state, make_images
from oracle import int_oracle as io
from util import check_reference_dets, int8_mismatch

pytestmark = pytest.mark.gpu
CFG = NetConfig(num_classes=20)


def _engine(calib, mode, res, max_batch):
    st = make_quant_state(CFG, calib, mode, res)
    return Engine.from_state_dict(CFG, st, res, res, max_batch, offset_mode=mode)


def _up2(a):
    return np.repeat(np.repeat(a, 2, axis=2), 2, axis=3)


@pytest.mark.parametrize("mode", ["round", "bilinear"])
def test_engine_matches_reference_vectors_256(golden, calib, mode):
    import torch
    g = golden("codenet1x_256_%s.npz" % mode)
    eng = _engine(calib, mode, 256, 4)
    x = make_images(2, 256, seed=2)
    eng.set_option("fuse_heads", 0)                     # the int8 grid between heads.dw2 and heads.out is only written unfused
    eng.set_option("fuse_units", 2)                     # the fused unit kernel also writes the two int8 grids inside every unit
    assert eng.units_fused == 11                        # 10 stride-1 units + branch 2 of layer1.0 (stride 2)
    out = eng.run(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    heads_unfused = eng.read_heads(2)
    # every int8 grid recorded from the reference, bit for bit
    names = {"hm.act1": ("heads.act1", slice(0, 64), True), "hm.act3": ("heads.act3", slice(0, 64), False)}
    checked = 0
    for k in g.files:
        if g[k].dtype != np.int8 or k.endswith(".s"):
            continue
        lbl, sl, up = names.get(k, (k, slice(None), False))
        got = eng.read_logical(lbl, 2)[:, sl]
        if up:
            got = _up2(got)
        assert int8_mismatch(got, g[k]) == 0, k
        checked += 1
    assert checked >= (16 if mode == "round" else 8)
    # the default path: heads.dw2 + heads.out as one kernel (heads_fused.cu); everything below is checked on ITS output
    eng.set_option("fuse_heads", 1)
    eng.set_option("fuse_units", 1)
    assert eng.heads_fused
    out = eng.run(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    # head outputs: exact integer accumulator, fp64 epilogue, one rounding to fp32
    heads = eng.read_heads(2)
    np.testing.assert_array_equal(heads, heads_unfused)
    ref = np.concatenate([g["hm_logit"], g["wh"], g["reg"]], 1)
    np.testing.assert_allclose(heads, ref, rtol=2e-7, atol=1e-7)
    hm = out["hm"].cpu().numpy()
    np.testing.assert_allclose(hm, 1 / (1 + np.exp(-g["hm_logit"])), rtol=2e-6, atol=1e-7)
    np.testing.assert_array_equal(out["wh"].cpu().numpy(), heads[:, 20:22])
    np.testing.assert_array_equal(out["reg"].cpu().numpy(), heads[:, 22:24])
    # detections: indices/order bit-exact against the deterministic oracle, tie-aware against the reference
    dets = out["dets"].cpu().numpy().astype(np.float64)
    inds = out["inds"].cpu().numpy()
    h64 = heads.astype(np.float64)
    odets, oinds = io.ctdet_decode(h64[:, :20], h64[:, 20:22], h64[:, 22:24], 100)
    np.testing.assert_array_equal(inds, oinds)
    np.testing.assert_allclose(dets, odets, rtol=1e-5, atol=1e-4)
    check_reference_dets(g["dets"][:2], h64[:, :20], h64[:, 20:22], h64[:, 22:24])
    eng.close()


DENSE_512 = ("stem", "layer1.out", "layer2.out", "layer3.out", "layer4", "up0.deform", "up0.out", "up1.deform", "up1.out",
             "up2.deform", "up2.out")


@pytest.mark.parametrize("mode", ["round", "bilinear"])
def test_engine_matches_reference_vectors_512(golden, calib, mode):
    """Config c geometry, two images, both offset modes: DENSE int8 grids of every stage output and of the whole deformable
    path, dense heads and the detections against the fp64 run of the unmodified reference (shufflenetv2_dcn.py:314-330)."""
    import torch
    g = golden("codenet1x_512_%s.npz" % mode)
    eng = _engine(calib, mode, 512, 2)
    x = make_images(2, 512, seed=3)
    out = eng.run(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    for k in DENSE_512:
        assert g[k].shape[0] == 2
        assert int8_mismatch(eng.read_logical(k, 2), g[k]) == 0, k
    heads = eng.read_heads(2)
    ref = np.concatenate([g["hm_logit"], g["wh"], g["reg"]], 1)          # fp32-rounded fp64 reference values
    np.testing.assert_allclose(heads, ref, rtol=2e-7, atol=1e-7)
    h64 = heads.astype(np.float64)
    odets, oinds = io.ctdet_decode(h64[:, :20], h64[:, 20:22], h64[:, 22:24], 100)
    np.testing.assert_array_equal(out["inds"].cpu().numpy(), oinds)
    np.testing.assert_allclose(out["dets"].cpu().numpy(), odets, rtol=1e-5, atol=1e-4)
    check_reference_dets(g["dets"][:2], h64[:, :20], h64[:, 20:22], h64[:, 22:24])
    eng.close()


def test_config_c_batch_256_equals_oracle(calib):
    """The BENCHMARKED configuration (BASELINE config c: 512^2, batch 256 on one GPU; bench.py's images: 16 distinct ones
    tiled): the last int8 grid, the heads and the top-K indices of the distinct images against the oracle, and every replica
    equal to its original -- through the graph path bench.py times and through the pipelined host path (submit / wait)."""
    import torch
    st = make_quant_state(CFG, calib, "round", 512)
    eng = Engine.from_state_dict(CFG, st, 512, 512, 256, offset_mode="round")
    base = make_images(16, 512, seed=100)
    x = np.concatenate([base] * 16)
    xt = torch.from_numpy(x).cuda()
    for _ in range(2):                                  # second call = graph replay, as in the timed loop
        out = eng.run(xt, maps=False)
    torch.cuda.synchronize()
    inds, dets = out["inds"].cpu().numpy(), out["dets"].cpu().numpy()
    up2 = eng.read_logical("up2.out", 256)
    heads = eng.read_heads(256)
    o = io.IntOracle(CFG, st, "round")
    for i in range(16):
        ref = o.forward(base[i:i + 1])
        assert int8_mismatch(up2[i:i + 1], o.cap["up2.out"]) == 0, i
        want = np.concatenate([ref["hm"], ref["wh"], ref["reg"]], 1).astype(np.float32)
        np.testing.assert_array_equal(heads[i:i + 1], want)
        h64 = want.astype(np.float64)
        odets, oinds = io.ctdet_decode(h64[:, :20], h64[:, 20:22], h64[:, 22:24], 100)
        np.testing.assert_array_equal(inds[i:i + 1], oinds)
        np.testing.assert_allclose(dets[i:i + 1], odets, rtol=1e-5, atol=1e-4)
    for r in range(1, 16):
        assert up2[16 * r:16 * r + 16].tobytes() == up2[:16].tobytes(), r
        assert heads[16 * r:16 * r + 16].tobytes() == heads[:16].tobytes(), r
        np.testing.assert_array_equal(inds[16 * r:16 * r + 16], inds[:16])
        np.testing.assert_array_equal(dets[16 * r:16 * r + 16], dets[:16])
    # pipelined host path: uint8 images (normalisation in the stem), two slots in flight, four steps
    mean, std = np.array([0.485, 0.456, 0.406], np.float32), np.array([0.229, 0.224, 0.225], np.float32)
    eng.set_normalization(mean, std)
    u8 = np.clip(np.rint((x.transpose(0, 2, 3, 1) * std + mean) * 255.0), 0, 255).astype(np.uint8)
    u8_t = torch.from_numpy(np.ascontiguousarray(u8)).pin_memory()
    want_d, want_i = eng.run_host(u8_t.numpy())
    bufs = [(torch.empty((256, 100, 6), dtype=torch.float32).pin_memory(), torch.empty((256, 100), dtype=torch.int32).pin_memory())
            for _ in range(2)]
    for step in range(4):
        sl = step & 1
        if step >= 2:
            eng.wait(sl)
            assert bufs[sl][1].numpy().tobytes() == want_i.tobytes() and bufs[sl][0].numpy().tobytes() == want_d.tobytes(), step
            bufs[sl][0].zero_(); bufs[sl][1].zero_()
        eng.submit_host(u8_t.numpy(), bufs[sl][0].numpy(), bufs[sl][1].numpy(), sl)
    with pytest.raises(_lib.CdnError):
        eng.submit_host(u8_t.numpy(), bufs[0][0].numpy(), bufs[0][1].numpy(), 0)          # slot 0 is still in flight
    for sl in (0, 1):
        eng.wait(sl)
        assert bufs[sl][1].numpy().tobytes() == want_i.tobytes() and bufs[sl][0].numpy().tobytes() == want_d.tobytes()
    eng.close()


def test_engine_equals_oracle_on_fresh_images(calib):
    """Other seeds than the golden ones (activations may saturate: the oracle saturates identically)."""
    import torch
    st = make_quant_state(CFG, calib, "round", 256)
    eng = Engine.from_state_dict(CFG, st, 256, 256, 4, offset_mode="round")
    x = make_images(3, 256, seed=77, clamp=4.0)
    out = eng.run(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    o = io.IntOracle(CFG, st, "round")
    ref = o.forward(x)
    n_int, n_guarded = eng.requant_stats          # every int8-producing layer of this network has an exact integer form
    assert n_guarded == 0 and n_int == sum(op.kind in ("dw", "deform") or (op.kind == "pw" and not op.a["n_f32"])
                                           for op in eng.plan.ops)
    for lbl in ("stem", "layer1.out", "layer2.out", "layer3.out", "layer4", "up0.deform", "up1.deform", "up2.deform", "up2.out"):
        assert int8_mismatch(eng.read_logical(lbl, 3), o.cap[lbl]) == 0, lbl
    heads = eng.read_heads(3)
    np.testing.assert_array_equal(heads, np.concatenate([ref["hm"], ref["wh"], ref["reg"]], 1).astype(np.float32))
    eng.close()


@pytest.mark.parametrize("res,batch", [(256, 5), (512, 3), (320, 2)])
def test_heads_fused_equals_separate_kernels(calib, res, batch):
    """heads.dw2 + heads.out as one kernel (heads_fused.cu: the depthwise conv produces the UMMA A tile in shared memory)
    against the two separate launches: fp32 head planes, indices and detections identical bit for bit, eager and as a
    graph, at tile counts that do and do not fill the persistent grid, and at a size (320 -> 80x80 maps) whose stored
    40x40 input is a whole number of 4x8 tiles only because 40 % 8 == 0."""
    import torch
    st = make_quant_state(CFG, calib, "round", 256)
    eng = Engine.from_state_dict(CFG, st, res, res, batch, offset_mode="round")
    xt = torch.from_numpy(make_images(batch, res, seed=11, clamp=4.0)).cuda()
    got = {}
    for fuse in (0, 1):
        eng.set_option("fuse_heads", fuse)
        assert eng.heads_fused == bool(fuse)
        for graph in (0, 1):
            eng.set_option("use_graph", graph)
            out = eng.run(xt, maps=False)
            torch.cuda.synchronize()
            got[fuse, graph] = (eng.read_heads(batch).copy(), out["inds"].cpu().numpy(), out["dets"].cpu().numpy())
    assert eng.num_launches > 0
    for key in ((0, 1), (1, 0), (1, 1)):
        for a, b in zip(got[0, 0], got[key]):
            np.testing.assert_array_equal(a, b)
    eng.close()


@pytest.mark.parametrize("res,batch", [(256, 5), (512, 3), (384, 2), (512, 40)])
def test_units_fused_equal_separate_kernels(calib, res, batch):
    """Every stride-1 ShuffleNetV2 unit of stages 2 and 3, and branch 2 of the first stride-2 unit (unit_s2_fused.cu: 1x1 conv on
    the 17 x 33 input pixels of a tile, stride-2 stencil), as ONE kernel each (unit_fused.cu: 1x1 conv on the halo'd tile -> int8
    tile in shared memory -> depthwise stencil -> A tile of the second 1x1 conv -> interleaving epilogue) against the three
    separate launches per unit: every tapped int8 grid -- the two tensors INSIDE each unit included, which the fused kernel
    writes only in its dump mode --, the heads and the detections identical bit for bit, eager and as a graph, at tile counts
    below and above the persistent grid (2 CTAs per SM) and at a size (384 -> 48x48 / 24x24 maps) where stage 3 is not
    eligible (24 % 16 != 0) and keeps its three launches."""
    import torch
    st = make_quant_state(CFG, calib, "round", 256)
    eng = Engine.from_state_dict(CFG, st, res, res, batch, offset_mode="round")
    xt = torch.from_numpy(make_images(batch, res, seed=13, clamp=4.0)).cuda()
    labels = [k for k in eng.plan.taps if k.startswith("layer")]
    inner = [k for k in labels if k.endswith("act1") or k.endswith("act2")]
    assert len(inner) >= 32

    def snapshot(names):
        out = eng.run(xt, maps=False)
        torch.cuda.synchronize()
        return ({k: eng.read_logical(k, batch) for k in names}, eng.read_heads(batch).copy(), out["inds"].cpu().numpy(),
                out["dets"].cpu().numpy())

    eng.set_option("fuse_units", 0)
    assert eng.units_fused == 0
    ref = snapshot(labels)
    n_launch0 = eng.num_launches
    eng.set_option("fuse_units", 2)
    n_fused = eng.units_fused
    assert n_fused == (4 if res == 384 else 11)         # + branch 2 of the first stride-2 unit (unit_s2_fused.cu)
    for graph in (0, 1):
        eng.set_option("use_graph", graph)
        got = snapshot(labels)
        for k in labels:
            assert int8_mismatch(got[0][k], ref[0][k]) == 0, (k, graph)
        for a, b in zip(got[1:], ref[1:]):
            np.testing.assert_array_equal(a, b)
    assert eng.num_launches == n_launch0 - 2 * n_fused
    eng.set_option("fuse_units", 1)                     # the default: nothing inside the units is written
    outer = [k for k in labels if k not in inner]
    # default kernels (warp-specialised for the wide stage), then the barrier-phased kernel everywhere (debug bit 29), then the
    # table-driven variants with the shifted requantisation (bit 22) -- all must give the same bits
    for flags in (0, 1 << 29, (1 << 29) | (1 << 22)):
        _lib.load().cdn_set_debug_flags(flags)
        eng.set_option("use_graph", 0)
        got = snapshot(outer)
        for k in outer:
            assert int8_mismatch(got[0][k], ref[0][k]) == 0, (k, flags)
        for a, b in zip(got[1:], ref[1:]):
            np.testing.assert_array_equal(a, b)
    _lib.load().cdn_set_debug_flags(0)
    eng.close()


def test_heads_fused_80_classes_equals_oracle(calib):
    """COCO-sized heads (80 + 2 + 2 planes: N = 96 UMMA columns, 128 TMEM columns, 2 resident CTAs per SM) through the fused
    heads kernel: bit-exact against the oracle and against the separate kernels.  Weights are synthetic for the 80-class
    geometry; BN statistics and ranges are borrowed from the 20-class calibration (activations may saturate: the oracle
    saturates identically)."""
    import torch
    cfg = NetConfig(num_classes=80)
    cal = {k: calib[k] for k in calib.files if k != "digest"}
    st = make_quant_state(cfg, cal, "round", 256)
    eng = Engine.from_state_dict(cfg, st, 256, 256, 2, offset_mode="round")
    x = make_images(2, 256, seed=21)
    xt = torch.from_numpy(x).cuda()
    assert eng.heads_fused
    out = eng.run(xt, maps=False)
    torch.cuda.synchronize()
    heads = eng.read_heads(2)
    ref = io.IntOracle(cfg, st, "round").forward(x)
    np.testing.assert_array_equal(heads, np.concatenate([ref["hm"], ref["wh"], ref["reg"]], 1).astype(np.float32))
    inds = out["inds"].cpu().numpy()
    eng.set_option("fuse_heads", 0)
    out0 = eng.run(xt, maps=False)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(eng.read_heads(2), heads)
    np.testing.assert_array_equal(out0["inds"].cpu().numpy(), inds)
    eng.close()


def test_batch_independence_graph_and_host_path(calib):
    """Frozen ranges make images independent: replicas give identical rows; graph replay, eager
    launches, the SIMT cross-check kernel and the host-buffer path all agree bit for bit."""
    import torch
    eng = _engine(calib, "round", 256, 8)
    base = make_images(2, 256, seed=5)
    x = np.concatenate([base, base, base[::-1], base])            # 8 images
    xt = torch.from_numpy(x).cuda()
    o1 = eng.run(xt, maps=False)
    torch.cuda.synchronize()
    d1, i1 = o1["dets"].cpu().numpy(), o1["inds"].cpu().numpy()
    np.testing.assert_array_equal(d1[0], d1[2]); np.testing.assert_array_equal(d1[0], d1[5]); np.testing.assert_array_equal(d1[1], d1[7])
    o2 = eng.run(xt, maps=False)                                   # second call = graph replay
    torch.cuda.synchronize()
    np.testing.assert_array_equal(o2["dets"].cpu().numpy(), d1)
    eng.set_option("use_graph", 0)
    o3 = eng.run(xt, maps=False)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(o3["dets"].cpu().numpy(), d1)
    _lib.load().cdn_set_debug_flags(1)
    o4 = eng.run(xt, maps=False)
    torch.cuda.synchronize()
    _lib.load().cdn_set_debug_flags(0)
    np.testing.assert_array_equal(o4["inds"].cpu().numpy(), i1)
    np.testing.assert_array_equal(o4["dets"].cpu().numpy(), d1)
    eng.set_option("use_graph", 1)
    eng.set_option("host_chunk", 3)
    hd, hi = eng.run_host(x)
    np.testing.assert_array_equal(hi, i1)
    np.testing.assert_array_equal(hd, d1)
    # a smaller batch through the same engine
    o5 = eng.run(xt[:3].contiguous(), maps=False)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(o5["dets"].cpu().numpy(), d1[:3])
    assert eng.num_launches >= 40                        # 44 with the fused units and heads tail (66 without)
    eng.close()


def test_config_c_geometry_large_batch(golden, calib):
    """BASELINE config c geometry (512^2) at a batch that exceeds L2: replicas must equal the single-image result,
    which is pinned to the reference vector."""
    import torch
    g = golden("codenet1x_512_round.npz")
    eng = _engine(calib, "round", 512, 64)
    two = make_images(2, 512, seed=3)
    x = np.concatenate([two] * 32)
    out = eng.run(torch.from_numpy(x).cuda(), maps=False)
    torch.cuda.synchronize()
    d = out["dets"].cpu().numpy()
    for b in range(2, 64):
        np.testing.assert_array_equal(d[b], d[b % 2])
    np.testing.assert_allclose(np.sort(d[0][:, 4]), np.sort(g["dets"][0][:, 4]), rtol=2e-6)
    eng.close()


def test_errors_are_loud(calib):
    import torch
    eng = _engine(calib, "round", 256, 2)
    with pytest.raises(_lib.CdnError):
        eng.run(torch.zeros((3, 3, 256, 256), device="cuda"))       # batch above the finalized maximum
    with pytest.raises(_lib.CdnError):
        eng.set_option("nonsense", 1)
    eng.close()


def test_engine_w2_maxpool_matches_reference_vectors(golden):
    """Config e geometry (2x width, stride-2 stem + MaxPool) against the fp64 run of the unmodified reference."""
    import torch
    from util import maxpool3s2_int8
    cfg = NetConfig(num_classes=20, w2=True, maxpool=True)
    g = golden("codenet_w2mp_256_round.npz")
    st = make_quant_state(cfg, golden("codenet_w2mp_calib.npz"), "round", 256)
    eng = Engine.from_state_dict(cfg, st, 256, 256, 2, offset_mode="round")
    x = make_images(2, 256, seed=2)[:1]
    eng.set_option("fuse_heads", 0)
    eng.set_option("fuse_units", 2)                     # stage-2 units (122 channels per half) run fused; dump the grids inside them
    assert eng.units_fused == 3
    out = eng.run(torch.from_numpy(x.copy()).cuda())
    torch.cuda.synchronize()
    heads_unfused = eng.read_heads(1)
    names = {"hm.act1": ("heads.act1", slice(0, 64), True), "hm.act3": ("heads.act3", slice(0, 64), False)}
    checked = 0
    for k in g.files:
        if g[k].dtype != np.int8 or k.endswith(".s"):
            continue
        lbl, sl, up = names.get(k, (k, slice(None), False))
        got = eng.read_logical(lbl, 1)[:, sl]
        if up:
            got = _up2(got)
        ref = maxpool3s2_int8(g[k]) if k == "stem" else g[k]
        assert int8_mismatch(got, ref) == 0, k
        checked += 1
    assert checked >= 17
    eng.set_option("fuse_heads", 1)
    assert eng.heads_fused
    out = eng.run(torch.from_numpy(x.copy()).cuda())
    torch.cuda.synchronize()
    heads = eng.read_heads(1)
    np.testing.assert_array_equal(heads, heads_unfused)
    np.testing.assert_allclose(heads, np.concatenate([g["hm_logit"], g["wh"], g["reg"]], 1), rtol=2e-7, atol=1e-7)
    h64 = heads.astype(np.float64)
    odets, oinds = io.ctdet_decode(h64[:, :20], h64[:, 20:22], h64[:, 22:24], 100)
    np.testing.assert_array_equal(out["inds"].cpu().numpy(), oinds)
    np.testing.assert_allclose(out["dets"].cpu().numpy(), odets, rtol=1e-5, atol=1e-4)
    eng.close()


def test_engine_w2_maxpool_512_matches_reference_vectors(golden):
    """BASELINE config e / config 4 geometry at its own resolution (w2, stride-2 stem + MaxPool, 512^2): dense int8 grids of
    the stage outputs and the deformable path (first deformable layer C = 2153), dense heads, detections."""
    import torch
    from util import maxpool3s2_int8
    cfg = NetConfig(num_classes=20, w2=True, maxpool=True)
    g = golden("codenet_w2mp_512_round.npz")
    st = make_quant_state(cfg, golden("codenet_w2mp_calib.npz"), "round", 512)
    eng = Engine.from_state_dict(cfg, st, 512, 512, 2, offset_mode="round")
    x = make_images(2, 512, seed=3)[:1]
    out = eng.run(torch.from_numpy(x.copy()).cuda())
    torch.cuda.synchronize()
    for k in DENSE_512:
        ref = maxpool3s2_int8(g[k]) if k == "stem" else g[k]
        assert int8_mismatch(eng.read_logical(k, 1), ref) == 0, k
    heads = eng.read_heads(1)
    np.testing.assert_allclose(heads, np.concatenate([g["hm_logit"], g["wh"], g["reg"]], 1), rtol=2e-7, atol=1e-7)
    h64 = heads.astype(np.float64)
    odets, oinds = io.ctdet_decode(h64[:, :20], h64[:, 20:22], h64[:, 22:24], 100)
    np.testing.assert_array_equal(out["inds"].cpu().numpy(), oinds)
    check_reference_dets(g["dets"][:1], h64[:, :20], h64[:, 20:22], h64[:, 22:24])
    eng.close()


def test_repeated_runs_are_identical_at_large_batch(calib):
    """Race / pipeline stress: 512x512, batch 96 (several tiles per SM in every GEMM, all epilogue groups and
    accumulator stages in rotation), 12 graph replays and the chunked host path must give identical bytes."""
    import torch
    st = make_quant_state(CFG, calib, "round", 512)
    eng = Engine.from_state_dict(CFG, st, 512, 512, 96, offset_mode="round")
    base = make_images(16, 512, seed=11)
    x = torch.from_numpy(np.concatenate([base] * 6)).cuda()
    ref = None
    for it in range(12):
        out = eng.run(x, maps=False, dets=True)
        torch.cuda.synchronize()
        cur = (out["dets"].cpu().numpy().tobytes(), out["inds"].cpu().numpy().tobytes(), eng.read_tensor(eng.plan.taps["up2.out"], 96).tobytes())
        if ref is None:
            ref = cur
        assert cur == ref, "run %d differs" % it
    # the 16 distinct images repeat 6 times: every copy must decode identically (batch independence)
    inds = np.frombuffer(ref[1], np.int32).reshape(96, -1)
    for r in range(1, 6):
        np.testing.assert_array_equal(inds[16 * r:16 * r + 16], inds[:16])
    d_host, i_host = eng.run_host(x.cpu().numpy())
    assert i_host.tobytes() == ref[1] and d_host.tobytes() == ref[0]
    eng.close()


def test_engine_from_plan_file_equals_engine_from_checkpoint(calib, tmp_path):
    """the compiled plan saved to disk (plan_io) and run without the checkpoint gives the same bytes."""
    import torch
    st = make_quant_state(CFG, calib, "bilinear", 256)
    eng = Engine.from_state_dict(CFG, st, 256, 256, 2, offset_mode="bilinear")
    f = str(tmp_path / "m.cdnplan.npz")
    eng.save_plan(f)
    x = torch.from_numpy(make_images(2, 256, seed=2)).cuda()
    a = eng.run(x)
    torch.cuda.synchronize()
    want = (eng.read_heads(2).copy(), a["dets"].cpu().numpy(), a["inds"].cpu().numpy())
    eng.close()
    eng2 = Engine.from_plan_file(f, 2)
    b = eng2.run(x)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(eng2.read_heads(2), want[0])
    np.testing.assert_array_equal(b["dets"].cpu().numpy(), want[1])
    np.testing.assert_array_equal(b["inds"].cpu().numpy(), want[2])
    eng2.close()
