"""Float (fp32) CoDeNet forward + decode on libcodenet_b200 -- BASELINE config 5 / SURVEY.md 8(a) a1, a10-a12.

Executes the UNQUANTISED network (`PoseShuffleNetV2.forward`, lib/models/networks/shufflenetv2_dcn.py:314-330) with
BatchNorm folded on the host (fp64, then rounded to fp32) and every layer as one of the fp32 NCHW kernels of
csrc/f32_net.cu / deform_f32.cu: conv + bias (+ ReLU), split / cat / channel_shuffle folded into channel-slice indexing,
the co-designed deformable module as ONE fused kernel (bilinear offsets, the reference's behaviour).  Results agree with
the reference evaluated in fp64 within 1e-4 relative (tests/test_gpu_f32.py).  This path is functional, not tuned:
launches are eager and the 1x1 convolutions are SIMT fp32 (TF32 tensor cores cannot hold the 1e-4 contract without a
3-way split); the optimised path of this repository is the W4A8 engine.

Input: a state dict in the reference's RAW (pre-quantisation) key space, e.g. `model.state_dict()` of the float model
or a `model_last.pth` of a float training run.  There is no CPU path.
"""
import ctypes as C
from typing import Dict

import numpy as np

from . import _lib
from .arch import NetConfig, build_graph

F = np.float64


def _fold(st, c, eps=1e-5):
    """conv (+ BatchNorm) -> (weight fp32, bias fp32): W' = W*gamma/sqrt(var+eps), b' = (b - mean)*gamma/sqrt(var+eps) + beta."""
    w = np.asarray(st[c.raw_conv + ".weight"], F)
    b = np.asarray(st[c.raw_conv + ".bias"], F) if c.has_bias else np.zeros(c.cout, F)
    if c.raw_bn:
        g, beta, mean, var = (np.asarray(st[c.raw_bn + "." + k], F) for k in ("weight", "bias", "running_mean", "running_var"))
        sf = g / np.sqrt(var + F(eps))
        w = w * sf.reshape(-1, 1, 1, 1)
        b = (b - mean) * sf + beta
    return w.astype(np.float32), b.astype(np.float32)


class EngineF32:
    def __init__(self, cfg: NetConfig, state: Dict[str, np.ndarray], device: int = 0, K: int = 100, gemm: str = "fp32"):
        import torch
        self.torch = torch
        self.cfg, self.K = cfg, int(K)
        self.lib = _lib.load()
        _lib.check(self.lib.cdn_check_device(device))
        self.dev = torch.device("cuda", device)
        self.g = build_graph(cfg)
        st = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in state.items()}
        assert gemm in ("tf32x3", "fp32"), gemm
        # 1x1 convs: "fp32" = SIMT fp32 products with fp64 block sums (tightest: half the reference's own fp32-vs-fp64 deviation);
        # "tf32x3" = tensor cores with a 3-way TF32 split (pw_tf32.cu): 1.7x the throughput of the whole model, deviation at the
        # level of the reference's own fp32 evaluation (tests/test_gpu_f32.py states both bounds)
        self.gemm = gemm
        self.P = {}
        self.Ptc = {}                           # name -> weights split and packed for cdn_pw_slice_tf32x3
        self.launches = 0                       # kernel launches issued so far (bench: gpu_launches per step)
        self.trace = None
        self._def_ws = None                     # scale scalars of the deformable modules (8 bytes per pixel), grown on demand
        self._dec_ws = None                     # ctdet decode candidate buffer, grown on demand
        self.scale_bias = {}                    # host copies of the offset-scale conv biases (kernel arguments by value)
        for c in self.g.all_convs():
            w, b = _fold(st, c)
            if c.kind == "scale":
                self.scale_bias[c.name] = float(b[0])
            self.P[c.name] = (torch.from_numpy(np.ascontiguousarray(w.reshape(c.cout, -1))).to(self.dev),
                              torch.from_numpy(np.ascontiguousarray(b)).to(self.dev))
        for kind in ("pw1", "dw2"):             # the heads' first two layers, stacked along the output channels
            self.P["heads." + kind] = tuple(torch.cat([self.P[h["name"] + "." + kind][i] for h in self.g.heads]).contiguous() for i in (0, 1))
        if gemm == "tf32x3":
            pw_names = [c.name for c in self.g.all_convs() if c.kind in ("pw", "head_out") and c.k == 1] + ["heads.pw1"]
            for name in pw_names:
                wd = self.P[name][0]
                n = int(self.lib.cdn_pw_tf32x3_packed_floats(wd.shape[0], wd.shape[1]))
                packed = torch.empty(n, dtype=torch.float32, device=self.dev)
                with torch.cuda.device(self.dev):
                    _lib.check(self.lib.cdn_pw_tf32x3_pack(self._p(wd), wd.shape[0], wd.shape[1], self._p(packed), self._st()))
                self.Ptc[name] = packed

    # -- thin kernel wrappers ---------------------------------------------------------------------------------------------
    def _st(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def _pw(self, x, in_off, cin, name, out, out_off, out_stride, relu):
        w, b = self.P[name]
        B, ct, H, W = x.shape
        self.launches += 1
        if self.trace is not None:                   # per-call CUDA events of the 1x1 convs (bench roofline), eager runs only
            e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            e0.record()
            self.trace, tr = None, self.trace
            self._pw(x, in_off, cin, name, out, out_off, out_stride, relu)
            self.launches -= 1
            self.trace = tr
            e1.record()
            tr.append((name, 2.0 * cin * w.shape[0] * B * H * W, e0, e1))
            return
        if name in self.Ptc and (H * W) % 256 == 0 and w.shape[1] == cin:
            _lib.check(self.lib.cdn_pw_slice_tf32x3(self._p(x), ct, in_off, cin, self._p(self.Ptc[name]), self._p(b), self._p(out), out.shape[1],
                                                    out_off, out_stride, w.shape[0], 1 if relu else 0, B, H * W, self._st()))
            return
        _lib.check(self.lib.cdn_pw_slice_f32(self._p(x), ct, in_off, cin, self._p(w), self._p(b), self._p(out), out.shape[1], out_off,
                                             out_stride, w.shape[0], 1 if relu else 0, B, H * W, self._st()))

    def _dw(self, x, name, stride, relu):
        w, b = self.P[name]
        self.launches += 1
        B, Cc, H, W = x.shape
        out = x.new_empty((B, Cc, (H - 1) // stride + 1, (W - 1) // stride + 1))
        _lib.check(self.lib.cdn_dw3x3_f32(self._p(x), self._p(w), self._p(b), self._p(out), B, Cc, H, W, stride, 1 if relu else 0, self._st()))
        return out

    def _new(self, x, c, H=None, W=None):
        return x.new_empty((x.shape[0], c, H if H else x.shape[2], W if W else x.shape[3]))

    # -- the network --------------------------------------------------------------------------------------------------------
    def forward(self, x):
        """x: CUDA fp32 [B,3,H,W] -> {'hm' (logits), 'wh', 'reg'} fp32 [B,*,H/4,W/4] (views of one heads tensor)."""
        torch, L, g, cfg = self.torch, self.lib, self.g, self.cfg
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
        x = x.contiguous()
        B = x.shape[0]
        with torch.cuda.device(self.dev):
            w, b = self.P["layer0"]
            s = g.stem.stride
            y = x.new_empty((B, g.stem.cout, (x.shape[2] - 1) // s + 1, (x.shape[3] - 1) // s + 1))
            self.launches += 1
            _lib.check(L.cdn_conv3x3_f32(self._p(x), self._p(w), self._p(b), self._p(y), B, 3, g.stem.cout, x.shape[2], x.shape[3], s, 1, self._st()))
            if cfg.maxpool:
                z = x.new_empty((B, y.shape[1], (y.shape[2] - 1) // 2 + 1, (y.shape[3] - 1) // 2 + 1))
                self.launches += 1
                _lib.check(L.cdn_maxpool3s2_f32(self._p(y), self._p(z), B * y.shape[1], y.shape[2], y.shape[3], self._st()))
                y = z
            x = y
            for u in g.units:                                      # BaseNode.forward, shufflenetv2_dcn.py:102-114
                r = "layer%d.%d." % (u["stage"], u["unit"])
                half = u["oup"] // 2
                if u["stride"] == 2:
                    d4 = self._dw(x, r + "dw4", 2, False)
                    out = self._new(d4, u["oup"])
                    self._pw(d4, 0, u["inp"], r + "pw5", out, 0, 2, True)              # x1 -> even channels
                    c1 = self._new(x, half)
                    self._pw(x, 0, u["inp"], r + "pw1", c1, 0, 1, True)
                    d2 = self._dw(c1, r + "dw2", 2, False)
                    self._pw(d2, 0, half, r + "pw3", out, 1, 2, True)                  # x2 -> odd channels
                else:
                    out = self._new(x, u["oup"])
                    self.launches += 1
                    _lib.check(L.cdn_copy_channels_f32(self._p(x), u["oup"], 0, self._p(out), u["oup"], 0, 2, half, B,
                                                       x.shape[2] * x.shape[3], self._st()))
                    c1 = self._new(x, half)
                    self._pw(x, half, half, r + "pw1", c1, 0, 1, True)
                    d2 = self._dw(c1, r + "dw2", 1, False)
                    self._pw(d2, 0, half, r + "pw3", out, 1, 2, True)
                x = out
            y = self._new(x, g.layer4.cout)
            self._pw(x, 0, g.layer4.cin, "layer4", y, 0, 1, True)
            x = y
            # up path: deform module + 1x1 conv + BN + ReLU + nearest x2.  No upsampled tensor is ever written: the next block's
            # deformable module reads its input through the virtual upsampling (cdn_deform_dw_up2_f32_ws, identical results), and
            # the last one is consumed the same way by the heads below.
            pending_up = False                                     # x still has to be read as its x2 upsampling
            z = None
            for up in g.ups:
                i = up["idx"]
                ws, bs = self.P["up%d.scale" % i]
                wd, _ = self.P["up%d.deform" % i]
                Bc, Cc, h, w = x.shape
                H, W = (2 * h, 2 * w) if pending_up else (h, w)
                y = x.new_empty((Bc, Cc, H, W))
                need = int(L.cdn_deform_dw_f32_ws_bytes(Bc, H, W, 1))
                if self._def_ws is None or self._def_ws.numel() < need:
                    self._def_ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
                self.launches += 2
                if pending_up:
                    _lib.check(L.cdn_deform_dw_up2_f32_ws(self._p(x), self._p(ws), C.c_float(self.scale_bias["up%d.scale" % i]), cfg.offset_bound,
                                                          self._p(wd), self._p(y), Bc, Cc, h, w, self._p(self._def_ws), self._def_ws.numel(), self._st()))
                else:
                    _lib.check(L.cdn_deform_dw_f32_ws(self._p(x), self._p(ws), C.c_float(self.scale_bias["up%d.scale" % i]), cfg.offset_bound,
                                                      self._p(wd), self._p(y), Bc, Cc, H, W, 1, self._p(self._def_ws), self._def_ws.numel(), self._st()))
                z = self._new(y, up["cout"])
                self._pw(y, 0, Cc, "up%d.channel" % i, z, 0, 1, True)
                x, pending_up = z, True
            if z is not None and z.shape[3] % 4 != 0:              # the heads' virtual upsampling needs rows of whole float4s
                Bc, Cc, h, w = z.shape
                x = z.new_empty((Bc, Cc, 2 * h, 2 * w))
                self.launches += 1
                _lib.check(L.cdn_upsample2x_f32(self._p(z), self._p(x), Bc * Cc, h, w, self._st()))
                z = None
            n_out = sum(h["classes"] for h in g.heads)
            off = 0
            views = {}
            # depthwise-separable heads, :244-271.  The three heads' first 1x1 conv and depthwise conv read the same tensor: they run
            # as ONE 64 -> 64*heads conv and ONE depthwise conv on the stacked channels (weights stacked at construction), each
            # head's output conv then reads its 64-channel slice -- the 128 x 128 feature map is read once instead of three times
            # A 1x1 conv + ReLU commutes with nearest upsampling exactly, so with the last upsampling still pending (z, low
            # resolution) the stacked conv runs on a quarter of the pixels and the depthwise conv reads its result through the virtual
            # upsampling (cdn_dw3x3_up2_f32): the upsampled 64-channel map and the upsampled conv output are never written.
            nh = len(g.heads)
            if z is not None:
                Bc, _, H, W = z.shape
                a = self._new(z, 64 * nh)
                self._pw(z, 0, 64, "heads.pw1", a, 0, 1, True)
                d = z.new_empty((Bc, 64 * nh, 2 * H, 2 * W))
                wdw, bdw = self.P["heads.dw2"]
                self.launches += 1
                _lib.check(L.cdn_dw3x3_up2_f32(self._p(a), self._p(wdw), self._p(bdw), self._p(d), Bc, 64 * nh, H, W, 1, self._st()))
            else:
                a = self._new(x, 64 * nh)
                self._pw(x, 0, 64, "heads.pw1", a, 0, 1, True)
                d = self._dw(a, "heads.dw2", 1, True)
            heads = self._new(d, n_out)
            for k, h in enumerate(g.heads):
                self._pw(d, 64 * k, 64, h["name"] + ".out", heads, off, 1, False)
                views[h["name"]] = heads[:, off:off + h["classes"]]
                off += h["classes"]
        self._heads = heads
        return views

    def detect(self, x, out=None):
        """forward + ctdet decode (lib/detectors/ctdet.py:31-41 without flip): dets [B,K,6], heads views.  The decode reads the
        head tensors as channel slices of the one [B, n_out, H, W] buffer the forward wrote (no copies) and takes its candidate
        buffer from a workspace kept between calls, so the whole call only enqueues kernels and can be captured in a CUDA graph
        (`out` = (dets, inds) to write into fixed buffers)."""
        torch = self.torch
        v = self.forward(x)
        hm, wh, reg = v["hm"], v["wh"], v.get("reg")
        B, cat, H, W = hm.shape
        for t in (hm, wh) + ((reg,) if reg is not None else ()):
            assert t.stride()[1:] == (H * W, W, 1), "head views must be channel slices of a contiguous tensor"
        dets, inds = out if out is not None else (torch.empty((B, self.K, 6), dtype=torch.float32, device=self.dev),
                                                  torch.empty((B, self.K), dtype=torch.int32, device=self.dev))
        need = int(self.lib.cdn_ctdet_decode_ws_bytes(B, cat, H, W))
        if self._dec_ws is None or self._dec_ws.numel() < need:
            self._dec_ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
        with torch.cuda.device(self.dev):
            self.launches += 2
            _lib.check(self.lib.cdn_ctdet_decode_ws(self._p(hm), hm.stride(0), self._p(wh), wh.stride(0), self._p(reg),
                                                    reg.stride(0) if reg is not None else 0, B, cat, H, W, self.K, 0, self._p(dets),
                                                    self._p(inds), self._p(self._dec_ws), self._dec_ws.numel(), self._st()))
        return dets, inds, v
