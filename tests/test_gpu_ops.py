"""GPU parity tests of the individual kernels, called through the C ABI (run on the B200 box with -m gpu)."""
import ctypes as C

import numpy as np
import pytest

import plan_sim
from codenet_b200 import _lib
from codenet_b200.arch import NetConfig
from codenet_b200.plan import build_plan, Plan, Op, TensorSpec
from codenet_b200.synth import make_quant_state, make_images
from oracle import int_oracle as io
from oracle import deform_ref
from util import int8_mismatch, assert_deform_f32_close

pytestmark = pytest.mark.gpu
CFG = NetConfig(num_classes=20)


@pytest.fixture(scope="module")
def sim256(calib):
    st = make_quant_state(CFG, calib, "round", 256)
    x = make_images(2, 256, seed=2)
    plan = build_plan(CFG, st, 256, 256, "round")
    T, heads = plan_sim.run_plan(plan, x)
    return plan, x, T, heads


@pytest.fixture(autouse=True)
def _reset_flags():
    _lib.load().cdn_set_debug_flags(0)
    yield
    _lib.load().cdn_set_debug_flags(0)


def test_device_is_sm100():
    _lib.check(_lib.load().cdn_check_device(0))


def _check_ops(sim256, kinds, flags=0):
    from gpu_util import run_op
    plan, x, T, heads = sim256
    _lib.load().cdn_set_debug_flags(flags)
    bad = []
    n = 0
    for op in plan.ops:
        if op.kind not in kinds:
            continue
        got = run_op(plan, op, T, images=x)
        n += 1
        if op.kind == "pw" and op.a["n_f32"]:
            if not np.array_equal(got, heads.astype(np.float32)):
                bad.append((op.name, float(np.abs(got - heads).max())))
            continue
        want = T[op.a["out_t"]]
        mism = int8_mismatch(got, want)
        if mism:
            bad.append((op.name, mism, want.size))
    assert n > 0
    assert not bad, "ops with mismatching elements: %s" % bad[:12]


def test_stem_every_layer(sim256):
    _check_ops(sim256, ("stem",))
    _check_ops(sim256, ("stem",), flags=1 << 26)        # the all-fp64 kernel (the definition), bit 26


def test_shuffle_unit_op_equals_three_ops(sim256):
    """cdn_shuffle_unit_i8 (QuantBaseNode.forward as one kernel: unit_fused.cu for the stride-1 units of stages 2 and 3,
    unit_s2_fused.cu for the main branch of the first stride-2 unit) against the simulated plan, unit by unit through the C ABI;
    units the fused kernels do not take (stage 4, the stride-2 units of stages 3 and 4) must answer "unit not fusable"."""
    import torch
    from codenet_b200 import ops
    plan, x, T, heads = sim256
    P = plan.ops
    fused, refused = [], []
    for i in range(len(P) - 2):
        a, b, c = P[i], P[i + 1], P[i + 2]
        if not (a.kind == "pw" and b.kind == "dw" and c.kind == "pw" and b.a["in_t"] == a.a["out_t"] and c.a["in_t"] == b.a["out_t"]
                and c.a["pass_t"] >= 0):
            continue
        bufs = {t: torch.from_numpy(T[t].astype(np.int8)).cuda() for t in (a.a["in_t"], c.a["pass_t"])}
        out = ops.run_unit(plan, a, b, c, bufs)
        if out is None:
            refused.append(a.name)
            continue
        assert int8_mismatch(out.cpu().numpy(), T[c.a["out_t"]]) == 0, a.name
        fused.append(a.name)
    assert len(fused) == 11 and fused[0] == "layer1.0.pw1", fused
    assert sorted(refused) == ["layer2.0.pw1", "layer3.0.pw1", "layer3.1.pw1", "layer3.2.pw1", "layer3.3.pw1"], refused


def test_stem_fast_path_equals_fp64_kernel(sim256):
    """stem.cu's guarded fp32 fast path (packed FFMA2 dot product, fp64 re-evaluation of every channel whose value is within the
    derived error bound of a rounding boundary) against the all-fp64 kernel on inputs chosen to stress the guard: ordinary images,
    huge and tiny magnitudes, exact cancellation between large taps, constant images whose requantised value sits ON a rounding
    boundary for some channel, and non-finite pixels.  Bit-identical int8 grids for both strides and with the max-pool."""
    from gpu_util import run_op
    plan, _, T, _ = sim256
    op = next(o for o in plan.ops if o.kind == "stem")
    rng = np.random.default_rng(7)
    H = W = 64
    imgs = [make_images(2, H, seed=9)]
    imgs.append((rng.standard_normal((2, 3, H, W)) * 1e6).astype(np.float32))
    imgs.append((rng.standard_normal((2, 3, H, W)) * 1e-6).astype(np.float32))
    big = rng.standard_normal((2, 3, H, W)).astype(np.float32) * 3e4
    big[:, :, ::2] = -big[:, :, 1::2]                                  # rows cancel pairwise: tiny sums of huge terms
    imgs.append(big)
    w = op.a["wq"].astype(np.float64).reshape(op.a["C"], 27)
    M, B = np.asarray(op.a["M"], np.float64), np.asarray(op.a["B"], np.float64)
    const = np.zeros((2, 3, H, W), np.float32)
    for b in range(2):                                                 # interior value v with M_c * v * sum(w_c) + B_c = k + 0.5 for channel c
        c = 3 + 5 * b
        k = np.floor(B[c]) + 2
        const[b] = np.float32((k + 0.5 - B[c]) / (M[c] * w[c].sum()))
    imgs.append(const)
    nf = make_images(2, H, seed=10)
    nf[0, 1, 10, 10] = np.nan; nf[1, 2, 20, 33] = np.inf; nf[1, 0, 5, 7] = -np.inf
    imgs.append(nf)
    L = _lib.load()
    for x in imgs:
        for stride, pool in ((4, 0), (2, 1), (2, 0)):
            o2 = Op("stem", "stem", dict(op.a, H=H, W=W, stride=stride, pool=pool))
            L.cdn_set_debug_flags(0)
            fast = run_op(plan, o2, T, images=x)
            L.cdn_set_debug_flags(1 << 26)
            ref = run_op(plan, o2, T, images=x)
            assert fast.shape == ref.shape and fast.tobytes() == ref.tobytes(), (stride, pool, int((fast != ref).sum()))
    L.cdn_set_debug_flags(0)


def test_depthwise_every_layer(sim256):
    _check_ops(sim256, ("dw",))


def test_deform_every_layer(sim256):
    _check_ops(sim256, ("deform",))


def test_pointwise_simt_every_layer(sim256):
    _check_ops(sim256, ("pw",), flags=1)


def test_pointwise_tcgen05_every_layer(sim256):
    _check_ops(sim256, ("pw",), flags=0)


def test_guarded_fp32_requant_fallback_every_layer(sim256):
    """Debug flag bit 7 builds the layers with the guarded fp32 requantisation (the path kept for channels that have
    no exact integer form): it must give the same bytes as the integer form the other tests run."""
    _check_ops(sim256, ("pw", "dw", "deform"), flags=128)


# ---- randomised shapes -------------------------------------------------------------------------------------------
def _mini_plan():
    return Plan(CFG, 0, 0, "round")


@pytest.mark.parametrize("C,H,W,stride,shift", [(24, 20, 28, 2, 0), (58, 17, 9, 1, 0), (116, 12, 12, 2, 0), (232, 8, 8, 1, 0),
                                                (192, 16, 16, 1, 1), (464, 6, 10, 1, 0), (2153, 4, 4, 1, 0)])
def test_depthwise_random(C, H, W, stride, shift):
    from gpu_util import run_op
    rng = np.random.default_rng(C * 7 + stride)
    P = _mini_plan()
    pitch = (C + 31) // 32 * 32
    tin = P.add_tensor(H >> shift, W >> shift, C, pitch)
    Hl, Wl = tin.H << shift, tin.W << shift
    tout = P.add_tensor((Hl - 1) // stride + 1, (Wl - 1) // stride + 1, C, pitch)
    wq = np.zeros((pitch, 9), np.int8); wq[:C] = rng.integers(-8, 8, (C, 9))
    M = np.zeros(pitch); B = np.zeros(pitch)
    M[:C] = rng.uniform(0.002, 0.02, C); B[:C] = rng.uniform(-140, -100, C)
    zx = int(rng.integers(100, 129))
    op = Op("dw", "t", dict(in_t=0, out_t=1, in_shift=shift, stride=stride, wq=wq, C=pitch, zx=zx, M=M, B=B, lo=-128 if C % 3 else -97))
    P.ops.append(op)
    x = np.zeros((3, tin.H, tin.W, pitch), np.int64); x[..., :C] = rng.integers(-128, 128, (3, tin.H, tin.W, C))
    sim = plan_sim.run_plan_seeded(P, {0: x})[0]
    got = run_op(P, op, {0: x})
    assert int8_mismatch(got, sim[1]) == 0
    assert len(np.unique(got)) > 20


@pytest.mark.parametrize("C,H,W,bound,shift,mode", [(128, 24, 24, 8, 0, 0), (256, 16, 16, 4, 1, 0), (1024, 8, 8, 8, 0, 0),
                                                    (58, 20, 12, 3, 0, 0), (24, 32, 32, 2, 0, 0), (464, 8, 8, 1, 0, 0),
                                                    (128, 12, 12, 8, 0, 1), (256, 8, 8, 4, 1, 1), (1024, 6, 6, 8, 0, 1),
                                                    (2153, 4, 4, 8, 0, 0)])
def test_deform_random(C, H, W, bound, shift, mode):
    """Config-2 style isolated layer: uniform s over the whole bound, int8 inputs, 4-bit weights, both modes."""
    from gpu_util import run_op
    rng = np.random.default_rng(C + 13 * bound + mode)
    P = _mini_plan()
    pitch = (C + 31) // 32 * 32
    tin = P.add_tensor(H >> shift, W >> shift, C, pitch)
    tout = P.add_tensor(H, W, C, pitch)
    Bn = 2
    wq = np.zeros((pitch, 9), np.int8); wq[:C] = rng.integers(-8, 8, (C, 9))
    M = np.zeros(pitch); Bc = np.zeros(pitch)
    M[:C] = rng.uniform(0.002, 0.01, C); Bc[:C] = rng.uniform(-20, 20, C)
    zx = 128
    ws = np.zeros(pitch, np.int8); ws[:C] = rng.integers(-8, 8, C)
    x = np.zeros((Bn, tin.H, tin.W, pitch), np.int64); x[..., :C] = rng.integers(-128, 128, (Bn, tin.H, tin.W, C))
    # choose Ms/bs so that u spans a bit more than [-bound+1, bound]
    acc_s = (x[..., :C] * ws[:C].astype(np.int64)).sum(-1) + zx * int(ws[:C].astype(np.int64).sum())
    span = max(1.0, float(np.abs(acc_s - acc_s.mean()).max()))
    Ms = (bound + 1.0) / span
    bs = 0.5 - Ms * float(acc_s.mean())
    ss, zs = io.act_params(-bound + 1, bound)
    lo = -90 if C in (58, 256) else -128               # lo > -128: the kernels' explicit lower clamp (ReLU with z_out != 128)
    a = dict(in_t=0, out_t=1, in_shift=shift, stride=1, wq=wq, C=pitch, zx=zx, M=M, B=Bc, lo=lo,
             ws=ws, Ms=Ms, bs=bs, ss=float(ss), zs=float(zs), bound=bound, mode=mode)
    op = Op("deform", "t", a)
    got, sval = run_op(P, op, {0: x}, sval=True)
    # oracle (NCHW) on the virtually upsampled input
    xin = x[..., :C]
    if shift:
        xin = np.repeat(np.repeat(xin, 2, axis=1), 2, axis=2)
    A = (xin + zx).transpose(0, 3, 1, 2)
    u = acc_s.astype(np.float64) * np.float64(Ms) + np.float64(bs)
    if shift:
        u = np.repeat(np.repeat(u, 2, axis=1), 2, axis=2)
    u = np.clip(u, -bound + 1.0, float(bound))
    qs = np.rint(ss * u - zs)
    s = (qs + zs) / ss
    wq3 = wq[:C].astype(np.int64).reshape(C, 3, 3)
    if mode == 0:
        s = np.rint(s)
        acc = io.deform_dw_int(A, wq3, s)
        q = np.clip(io.requant(acc, M[:C], Bc[:C], 0, False), lo, 127)
    else:
        acc = io.deform_dw_bilinear(A, wq3, s)
        q = np.clip(io.requant_f(acc, M[:C], Bc[:C], 0, False), lo, 127)
    np.testing.assert_array_equal(sval, s.astype(np.float32))
    assert len(np.unique(s)) >= min(4, 2 * bound)
    assert int8_mismatch(got[..., :C].transpose(0, 3, 1, 2), q) == 0
    assert not got[..., C:].any()


@pytest.mark.parametrize("K,N,relu,flags", [(24, 58, 1, 0), (58, 58, 1, 0), (116, 232, 0, 0), (464, 1024, 1, 0), (1024, 256, 1, 0),
                                            (64, 192, 1, 0), (200, 64, 0, 0), (58, 58, 1, 1), (464, 1024, 1, 1),
                                            (58, 58, 1, 128), (464, 1024, 0, 128)])
def test_pointwise_random_dense(K, N, relu, flags):
    from gpu_util import run_op
    _lib.load().cdn_set_debug_flags(flags)
    rng = np.random.default_rng(K * 31 + N)
    P = _mini_plan()
    Kp, Np = (K + 31) // 32 * 32, (N + 31) // 32 * 32
    H, W, Bn = 9, 15, 3                                                # 405 pixels: a partial last tile
    tin = P.add_tensor(H, W, K, Kp)
    tout = P.add_tensor(H, W, N, Np)
    N16 = (N + 15) // 16 * 16
    w = np.zeros((N16, Kp), np.int8); w[:N, :K] = rng.integers(-8, 8, (N, K))
    M = np.zeros(N16); Bc = np.zeros(N16)
    M[:N] = rng.uniform(0.2, 1.0, N) / np.sqrt(K) / 30; Bc[:N] = rng.uniform(-130, -90, N)
    chunks = [((16 * j) if N - 16 * j > 0 else 0, int(np.clip(N - 16 * j, 0, 16)), -1, 16 * j) for j in range(Np // 16)]
    a = dict(in_t=0, out_t=1, pass_t=-1, k_off=0, K=Kp, N=N16, zx=128, wq=w, M=M, B=Bc, lo=-128 if not relu else -101,
             chunks=np.array(chunks, np.int16), n_f32=0)
    op = Op("pw", "t", a)
    P.ops.append(op)
    x = np.zeros((Bn, H, W, Kp), np.int64); x[..., :K] = rng.integers(-128, 128, (Bn, H, W, K))
    sim = plan_sim.run_plan_seeded(P, {0: x})[0]
    got = run_op(P, op, {0: x})
    assert int8_mismatch(got, sim[1]) == 0
    assert len(np.unique(got)) > (20 if not relu else 8)


def test_general_deform_conv_f32(golden):
    """cdn_deform_conv_forward_f32 against the vectors of the reference's op (B1 boundary)."""
    import torch
    from gpu_util import dev, ptr, stream
    g = golden("deform_kat.npz")
    L = _lib.load()
    for name in ("dw_s1", "dw_s2", "dense", "g2"):
        x, off, w, y = (g[name + s] for s in ("_x", "_off", "_w", "_y"))
        st, pad, dil, groups, dg = (int(v) for v in g[name + "_cfg"])
        Bn, Cc, H, W = x.shape
        Co = w.shape[0]
        out = torch.zeros(y.shape, dtype=torch.float32, device="cuda")
        tx, to, tw = dev(x.astype(np.float32)), dev(off.astype(np.float32)), dev(w.astype(np.float32))
        _lib.check(L.cdn_deform_conv_forward_f32(ptr(tx), ptr(tw), ptr(to), ptr(out), Bn, Cc, H, W, Co, 3, 3, st, st,
                                                 pad, pad, dil, dil, groups, dg, 64, stream()))
        torch.cuda.synchronize()
        # every element within 1e-4 relative; floor()-boundary flips (if any) explicitly bounded, none ignored
        assert_deform_f32_close(out.cpu().numpy(), x, off, w, y, st, pad, dil, groups, dg, name)
    # error behaviour of shape_check (dcn_deform_conv_cuda.cpp:61-149)
    assert L.cdn_deform_conv_forward_f32(ptr(tx), ptr(tw), ptr(to), ptr(out), 1, 4, 2, 2, 4, 3, 3, 1, 1, 0, 0, 1, 1, 1, 1, 64, stream()) == -1


@pytest.mark.parametrize("name", ["voc", "small"])
def test_decode_kat_gpu(golden, name):
    import torch
    from gpu_util import dev, ptr, stream
    g = golden("decode_kat.npz")
    hm = g[name + "_hm"]
    logit = (np.log(hm.astype(np.float64)) - np.log1p(-hm.astype(np.float64))).astype(np.float32)
    Bn, cat, H, W = hm.shape
    K = int(g[name + "_K"])
    L = _lib.load()
    for reg, ref in ((g[name + "_reg"], g[name + "_dets"]), (None, g[name + "_dets_noreg"])):
        dets = torch.zeros((Bn, K, 6), dtype=torch.float32, device="cuda")
        inds = torch.zeros((Bn, K), dtype=torch.int32, device="cuda")
        t_hm, t_wh, t_reg = dev(logit), dev(g[name + "_wh"]), (dev(reg) if reg is not None else None)   # keep alive
        _lib.check(L.cdn_ctdet_decode(ptr(t_hm), ptr(t_wh), ptr(t_reg), Bn, cat, H, W, K, ptr(dets), ptr(inds), stream()))
        d = dets.cpu().numpy()
        odets, oinds = io.ctdet_decode(logit.astype(np.float64), g[name + "_wh"].astype(np.float64),
                                       None if reg is None else reg.astype(np.float64), K)
        np.testing.assert_array_equal(inds.cpu().numpy(), oinds)                    # bit-exact top-K indices and order
        np.testing.assert_allclose(d[..., :4], ref[..., :4], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(d[..., 4], ref[..., 4], rtol=2e-6)
        np.testing.assert_array_equal(d[..., 5], ref[..., 5])


def test_decode_workspace_form_on_channel_slices():
    """cdn_ctdet_decode_ws on three channel slices of one [B, C, H, W] head tensor (image strides, caller's workspace, no stream
    wait) = cdn_ctdet_decode on contiguous copies, bit for bit; too small a workspace is refused."""
    import ctypes as C
    import torch
    from gpu_util import ptr, stream
    rng = np.random.default_rng(5)
    Bn, cat, H, W, K = 4, 20, 64, 64, 100
    heads = torch.from_numpy(rng.normal(-2, 2, (Bn, cat + 4, H, W)).astype(np.float32)).cuda()
    hm, wh, reg = heads[:, :cat], heads[:, cat:cat + 2], heads[:, cat + 2:]
    L = _lib.load()
    want_d, want_i = torch.zeros((Bn, K, 6), device="cuda"), torch.zeros((Bn, K), dtype=torch.int32, device="cuda")
    hc, wc, rc = hm.contiguous(), wh.contiguous(), reg.contiguous()
    _lib.check(L.cdn_ctdet_decode(ptr(hc), ptr(wc), ptr(rc), Bn, cat, H, W, K, ptr(want_d), ptr(want_i), stream()))
    need = int(L.cdn_ctdet_decode_ws_bytes(Bn, cat, H, W))
    assert need == Bn * cat * H * W * 8 + (Bn + 1) * 4
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    for it in range(2):                                               # second call: the workspace is re-used as it was left
        got_d, got_i = torch.zeros_like(want_d), torch.zeros_like(want_i)
        _lib.check(L.cdn_ctdet_decode_ws(ptr(hm), hm.stride(0), ptr(wh), wh.stride(0), ptr(reg), reg.stride(0), Bn, cat, H, W, K, 0,
                                         ptr(got_d), ptr(got_i), ptr(ws), need, stream()))
        torch.cuda.synchronize()
        assert torch.equal(got_i, want_i) and torch.equal(got_d, want_d)
    assert L.cdn_ctdet_decode_ws(ptr(hm), hm.stride(0), ptr(wh), wh.stride(0), ptr(reg), reg.stride(0), Bn, cat, H, W, K, 0,
                                 ptr(got_d), ptr(got_i), ptr(ws), need - 8, stream()) == -1
    assert b"workspace" in L.cdn_last_error()


@pytest.mark.parametrize("cat,H,W,K,levels", [(20, 64, 64, 100, 37), (80, 32, 48, 100, 500), (3, 16, 16, 40, 5), (1, 8, 8, 100, 3),
                                              (20, 128, 128, 100, 2000)])
def test_decode_with_ties(cat, H, W, K, levels):
    """Quantised heads produce heavy ties; order = (logit desc, class asc, index asc), fewer-than-K peaks -> -1."""
    import torch
    from gpu_util import dev, ptr, stream
    rng = np.random.default_rng(cat + H + levels)
    Bn = 3
    logit = (rng.integers(0, levels, (Bn, cat, H, W)) / levels * 8 - 6).astype(np.float32)
    wh = rng.uniform(1, 10, (Bn, 2, H, W)).astype(np.float32)
    reg = rng.uniform(0, 1, (Bn, 2, H, W)).astype(np.float32)
    dets = torch.zeros((Bn, K, 6), dtype=torch.float32, device="cuda")
    inds = torch.zeros((Bn, K), dtype=torch.int32, device="cuda")
    t_hm, t_wh, t_reg = dev(logit), dev(wh), dev(reg)                  # keep the device buffers alive
    _lib.check(_lib.load().cdn_ctdet_decode(ptr(t_hm), ptr(t_wh), ptr(t_reg), Bn, cat, H, W, K, ptr(dets),
                                            ptr(inds), stream()))
    got = inds.cpu().numpy()
    for b in range(Bn):
        p = np.full((cat, H + 2, W + 2), -np.inf, np.float32); p[:, 1:-1, 1:-1] = logit[b]
        mx = np.max([p[:, i:i + H, j:j + W] for i in range(3) for j in range(3)], axis=0)
        flat = np.where(mx == logit[b], logit[b], -np.inf).reshape(-1)
        order = np.lexsort((np.arange(flat.size), -flat))
        npk = int(np.isfinite(flat).sum())
        want = np.full(K, -1, np.int64); want[:min(K, npk)] = order[:min(K, npk)]
        np.testing.assert_array_equal(got[b], want)
    d = dets.cpu().numpy()
    assert np.all(d[got < 0] == 0)
