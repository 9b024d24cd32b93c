"""Python host of the engine: feeds a Plan to libcodenet_b200 through the C ABI and runs it.

PyTorch is used only for device memory and streams (plumbing); every kernel is hand-written CUDA in csrc/.
There is no CPU path: constructing an Engine without an sm_100 GPU raises CdnError.
"""
import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

from . import _lib
from .arch import NetConfig
from .plan import Plan, build_plan


class Engine:
    def __init__(self, plan: Plan, max_batch: int, device: int = 0, K: int = 100):
        self.plan, self.max_batch, self.device, self.K = plan, int(max_batch), int(device), int(K)
        self.lib = _lib.load()
        if os.environ.get("CODENET_DEBUG_FLAGS"):          # kernel experiments only
            self.lib.cdn_set_debug_flags(int(os.environ["CODENET_DEBUG_FLAGS"]))
        if os.environ.get("CODENET_PW_SIMT") == "1":      # bring-up switch: SIMT cross-check kernel for 1x1 convs
            self.lib.cdn_set_debug_flags(1)
        self._h = C.c_void_p()
        _lib.check(self.lib.cdn_engine_create(C.byref(self._h), self.device))
        keep = _lib.Keep()
        L = self.lib
        for t in plan.tensors:
            tid = L.cdn_engine_add_tensor(self._h, t.H, t.W, t.pitch)
            if tid != t.id:
                _lib.check(tid if tid < 0 else -1)
        for op in plan.ops:
            a = op.a
            if op.kind == "stem":
                rq = keep.requant(a["M"], a["B"], a["lo"])
                rc = L.cdn_engine_add_stem(self._h, a["out_t"], a["H"], a["W"], a["stride"], a["pool"], keep.i8(a["wq"]),
                                           a["C"], C.byref(rq))
            elif op.kind == "dw":
                rq = keep.requant(a["M"], a["B"], a["lo"])
                rc = L.cdn_engine_add_dw(self._h, a["in_t"], a["out_t"], a["in_shift"], a["stride"], keep.i8(a["wq"]),
                                         a["C"], a["zx"], C.byref(rq))
            elif op.kind == "deform":
                rq = keep.requant(a["M"], a["B"], a["lo"])
                sc = keep.deform_scale(a)
                rc = L.cdn_engine_add_deform(self._h, a["in_t"], a["out_t"], a["in_shift"], C.byref(sc), keep.i8(a["wq"]),
                                             a["C"], a["zx"], C.byref(rq))
            elif op.kind == "pw":
                d = keep.pw_desc(a)
                rc = L.cdn_engine_add_pw(self._h, a["in_t"], a["pass_t"], a["out_t"], C.byref(d))
            else:
                raise ValueError(op.kind)
            if rc != 0:
                raise _lib.CdnError("codenet_b200: op %s: %s" % (op.name, L.cdn_last_error().decode()))
        has_reg = 1 if len(plan.cfg.head_list()) > 2 else 0
        _lib.check(L.cdn_engine_set_heads(self._h, plan.cat, plan.out_H, plan.out_W, self.K, has_reg))
        _lib.check(L.cdn_engine_finalize(self._h, self.max_batch))
        self.n_f32 = plan.cat + 2 + 2 * has_reg

    # -- construction helpers -------------------------------------------------------------------------------------
    @classmethod
    def from_state_dict(cls, cfg: NetConfig, state: Dict[str, np.ndarray], in_H: int, in_W: int, max_batch: int,
                        offset_mode: str = "bilinear", device: int = 0, K: int = 100):
        """`offset_mode`: "bilinear" = the reference's semantics (fractional offsets, quant_modules.py:668-671; the default
        everywhere); "round" = the co-designed integer-offset mode (explicit opt-in)."""
        state = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in state.items()}
        return cls(build_plan(cfg, state, in_H, in_W, offset_mode), max_batch, device, K)

    @classmethod
    def from_module(cls, model, in_H: int, in_W: int, max_batch: int, offset_mode: str = "bilinear", device: int = 0,
                    K: int = 100):
        """`model`: a PoseShuffleNetV2 already rewritten by quantize_shufflenetv2_dcn (reference or our mirror)."""
        sd = model.state_dict()
        w2 = sd["layer4.0.conv.weight"].shape[0] == 2153
        maxpool = tuple(getattr(model.layer0[0].conv, "stride", (4, 4))) == (2, 2)
        heads = tuple((h, int(sd[h + ".quant_conv.weight"].shape[0])) for h in model.heads)
        bound = 8
        try:                                  # Hardtanh(-bound+1, bound) of the first deformable block
            bound = int(round(float(model.deconv_layers[0].quant_act[0].max_val)))
        except Exception:
            pass
        cfg = NetConfig(num_classes=heads[0][1], w2=w2, maxpool=maxpool, heads=heads, offset_bound=bound,
                        wt_percentile=bool(getattr(model, "wt_percentile", False)))
        return cls.from_state_dict(cfg, sd, in_H, in_W, max_batch, offset_mode, device, K)

    @classmethod
    def from_plan_file(cls, path: str, max_batch: int, device: int = 0, K: int = 100):
        """Runs a compiled plan saved by codenet_b200.plan_io.save_plan: no checkpoint, network definition or torch.load."""
        from .plan_io import load_plan
        return cls(load_plan(path), max_batch, device, K)

    def save_plan(self, path: str):
        from .plan_io import save_plan
        save_plan(self.plan, path)

    def close(self):
        if self._h:
            self.lib.cdn_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int):
        _lib.check(self.lib.cdn_engine_set_option(self._h, name.encode(), int(value)))

    # -- execution ----------------------------------------------------------------------------------------------------
    def set_normalization(self, mean, std):
        """mean / std as the reference holds them (float32 [3], lib/opts.py): enables uint8 input."""
        m = np.ascontiguousarray(np.asarray(mean, np.float32).reshape(3))
        s = np.ascontiguousarray(np.asarray(std, np.float32).reshape(3))
        _lib.check(self.lib.cdn_engine_set_normalization(self._h, C.c_void_p(m.ctypes.data), C.c_void_p(s.ctypes.data)))
        self._norm = True

    def run(self, images, maps: bool = True, dets: bool = True, out: Optional[dict] = None, raw_hm: bool = False):
        """images: torch CUDA tensor (contiguous), either fp32 [B,3,H,W] (normalised, what the reference's model takes) or
        uint8 [B,H,W,3] (what cv2 hands to pre_process; needs set_normalization).  Returns dict of torch CUDA tensors:
        hm (post-sigmoid, ctdet.py:32; logits when raw_hm), wh, reg [B,*,H/4,W/4]; dets [B,K,6]; inds [B,K]."""
        import torch
        if raw_hm != getattr(self, "_raw_hm", False):
            self.set_option("hm_logits", 1 if raw_hm else 0)
            self._raw_hm = raw_hm
        assert images.is_cuda and images.is_contiguous()
        u8 = images.dtype == torch.uint8
        B = images.shape[0]
        if u8:
            assert images.shape[1:] == (self.plan.in_H, self.plan.in_W, 3), images.shape
        else:
            assert images.dtype == torch.float32 and images.shape[1:] == (3, self.plan.in_H, self.plan.in_W), images.shape
        Ho, Wo, cat = self.plan.out_H, self.plan.out_W, self.plan.cat
        o = out if out is not None else {}
        dev = images.device
        if maps and "hm" not in o:
            o["hm"] = torch.empty((B, cat, Ho, Wo), dtype=torch.float32, device=dev)
            o["wh"] = torch.empty((B, 2, Ho, Wo), dtype=torch.float32, device=dev)
            o["reg"] = torch.empty((B, 2, Ho, Wo), dtype=torch.float32, device=dev)
        if dets and "dets" not in o:
            o["dets"] = torch.empty((B, self.K, 6), dtype=torch.float32, device=dev)
            o["inds"] = torch.empty((B, self.K), dtype=torch.int32, device=dev)
        p = lambda k: C.c_void_p(o[k].data_ptr()) if k in o else None
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        fn = self.lib.cdn_engine_run_u8 if u8 else self.lib.cdn_engine_run
        _lib.check(fn(self._h, C.c_void_p(images.data_ptr()), B, p("hm") if maps else None,
                      p("wh") if maps else None, p("reg") if maps else None,
                      p("dets") if dets else None, p("inds") if dets else None, stream))
        return o

    def run_host(self, images: np.ndarray, dets: Optional[np.ndarray] = None, inds: Optional[np.ndarray] = None):
        """images: host array (ideally backed by pinned memory), fp32 [B,3,H,W] or uint8 [B,H,W,3].
        H2D (chunked, overlapped with compute), forward, decode, D2H of the detections."""
        B = images.shape[0]
        assert images.dtype in (np.float32, np.uint8) and images.flags["C_CONTIGUOUS"]
        if dets is None:
            dets = np.empty((B, self.K, 6), np.float32)
        if inds is None:
            inds = np.empty((B, self.K), np.int32)
        fn = self.lib.cdn_engine_run_host_u8 if images.dtype == np.uint8 else self.lib.cdn_engine_run_host
        _lib.check(fn(self._h, C.c_void_p(images.ctypes.data), B, C.c_void_p(dets.ctypes.data), C.c_void_p(inds.ctypes.data)))
        return dets, inds

    def submit_host(self, images: np.ndarray, dets: np.ndarray, inds: Optional[np.ndarray], slot: int):
        """Pipelined run_host: enqueue H2D -> forward + decode -> D2H for `slot` (0 / 1) and return at once; `wait(slot)`
        blocks until `dets` / `inds` hold the step's result.  Keep both slots in flight and the copy of step n+1 overlaps
        the compute of step n.  All three host arrays must be pinned and stay alive until `wait`."""
        assert images.dtype in (np.float32, np.uint8) and images.flags["C_CONTIGUOUS"]
        fn = self.lib.cdn_engine_submit_host_u8 if images.dtype == np.uint8 else self.lib.cdn_engine_submit_host
        _lib.check(fn(self._h, C.c_void_p(images.ctypes.data), images.shape[0], C.c_void_p(dets.ctypes.data),
                      C.c_void_p(inds.ctypes.data) if inds is not None else None, int(slot)))

    def wait(self, slot: int):
        _lib.check(self.lib.cdn_engine_wait(self._h, int(slot)))

    def profile(self, images):
        """Per-op device times (ms) of one eager run: list of (op name, kind, ms) + ('decode', ...)."""
        import torch
        n = len(self.plan.ops) + 2
        ms = np.zeros(n, np.float32)
        stream = C.c_void_p(torch.cuda.current_stream(images.device).cuda_stream)
        _lib.check(self.lib.cdn_engine_profile(self._h, C.c_void_p(images.data_ptr()), images.shape[0],
                                               C.c_void_p(ms.ctypes.data), n, stream))
        out = [(op.name, op.kind, float(ms[i])) for i, op in enumerate(self.plan.ops)]
        out.append(("heads_copyout", "copyout", float(ms[n - 2])))
        out.append(("ctdet_decode", "decode", float(ms[n - 1])))
        return out

    @property
    def heads_fused(self) -> bool:
        """True when heads.dw2 + heads.out run as one kernel (option "fuse_heads", default on, and the pair is eligible)."""
        return bool(self.lib.cdn_engine_heads_fused(self._h))

    @property
    def units_fused(self) -> int:
        """How many stride-1 ShuffleNetV2 units run as one kernel (option "fuse_units", default on; unit_fused.cu)."""
        return int(self.lib.cdn_engine_units_fused(self._h))

    def op_fusion(self, i: int) -> int:
        """How plan op i runs: 0 = its own launch, 1 = head of the fused heads tail, 2 = head of a fused unit, -1 = folded into an
        earlier op's launch."""
        return int(self.lib.cdn_engine_op_fusion(self._h, int(i)))

    @property
    def num_launches(self):
        return int(self.lib.cdn_engine_num_launches(self._h))

    @property
    def requant_stats(self):
        """(layers on the exact integer requantisation, layers on the guarded fp32 sequence)."""
        a, b = C.c_int(0), C.c_int(0)
        _lib.check(self.lib.cdn_engine_requant_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- test / debug access ------------------------------------------------------------------------------------------
    def read_tensor(self, tid: int, batch: int) -> np.ndarray:
        t = self.plan.tensors[tid]
        buf = np.empty((batch, t.H, t.W, t.pitch), np.int8)
        _lib.check(self.lib.cdn_engine_read_tensor(self._h, tid, batch, C.c_void_p(buf.ctypes.data)))
        return buf

    def read_logical(self, label: str, batch: int) -> np.ndarray:
        """Activation of a tapped QuantAct as logical NCHW int8."""
        t = self.plan.tensors[self.plan.taps[label]]
        raw = self.read_tensor(t.id, batch)
        return np.ascontiguousarray(raw[..., t.phys(np.arange(t.C))].transpose(0, 3, 1, 2))

    def read_heads(self, batch: int) -> np.ndarray:
        buf = np.empty((batch, self.n_f32, self.plan.out_H, self.plan.out_W), np.float32)
        _lib.check(self.lib.cdn_engine_read_heads(self._h, batch, C.c_void_p(buf.ctypes.data)))
        return buf
