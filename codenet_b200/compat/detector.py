"""Detector entry points mirror: lib/detectors/base_detector.py:20-155 and lib/detectors/ctdet.py:26-74.

`CtdetDetector(opt)` builds the network, quantises it, loads the checkpoint (`opt.load_model`, reference key space),
freezes the activation ranges and runs forward + decode through the compiled int8 engine.  `run()` returns the same
dict as the reference (`results` keyed by 1-based class id and the stage timers test.py prints, test.py:76-79).
Only what differs from the reference lives here: engine construction, process / process_u8, the device-side box transform.
pre_process, merge_outputs and run keep the reference's contract (same inputs, outputs and timer keys) in this package's own
decomposition (input_geometry / _Stages / group_by_class); INTEGRATION.md section 2 shows the patch for users who keep the reference's BaseDetector.
"""
import time
from types import SimpleNamespace

import numpy as np
import torch

from .nms import soft_nms
from .decode import ctdet_decode
from .quantize_model import quantize_shufflenetv2_dcn, freeze_ranges
from .shufflenetv2_dcn import PoseShuffleNetV2


def default_opt(**kw):
    """The fields of lib/opts.py that reach the hot path, with the reference's defaults for ctdet / shufflenetv2."""
    o = SimpleNamespace(
        gpus=[0], arch="shufflenetv2", heads={"hm": 20, "wh": 2, "reg": 2}, head_conv=64, num_classes=20,
        resume_quantize=True, w_bit=4, a_bit=8, wt_percentile=False, act_percentile=False, w2=False, maxpool=False,
        load_model="", mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225], test_scales=[1.0], fix_res=True,
        input_h=512, input_w=512, pad=31, down_ratio=4, flip_test=False, reg_offset=True, cat_spec_wh=False, K=100,
        nms=False, debug=0, offset_mode="bilinear", max_batch=1)
    o.__dict__.update(kw)
    return o


# ---- affine helpers (lib/utils/image.py:14-61) restated with numpy only --------------------------------------------
def _third_point(a, b):
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def _solve_affine(src, dst):
    """2x3 matrix mapping three src points to three dst points (what cv2.getAffineTransform computes)."""
    A = np.hstack([src.astype(np.float64), np.ones((3, 1))])
    return np.linalg.solve(A, dst.astype(np.float64)).T


def get_affine_transform(center, scale, rot, output_size, shift=np.array([0, 0], dtype=np.float32), inv=0):
    if not isinstance(scale, (np.ndarray, list)):
        scale = np.array([scale, scale], dtype=np.float32)
    scale = np.asarray(scale, dtype=np.float32)
    src_w, dst_w, dst_h = scale[0], output_size[0], output_size[1]
    rad = np.pi * rot / 180
    sn, cs = np.sin(rad), np.cos(rad)
    src_dir = np.array([-(src_w * -0.5) * sn, (src_w * -0.5) * cs], dtype=np.float32)
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), np.float32); dst = np.zeros((3, 2), np.float32)
    src[0] = center + scale * shift
    src[1] = center + src_dir + scale * shift
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5], np.float32) + dst_dir
    src[2] = _third_point(src[0], src[1]); dst[2] = _third_point(dst[0], dst[1])
    return _solve_affine(dst, src) if inv else _solve_affine(src, dst)


def transform_preds(coords, center, scale, output_size):
    t = get_affine_transform(center, scale, 0, output_size, inv=1)
    pts = np.concatenate([coords[:, 0:2].astype(np.float32), np.ones((coords.shape[0], 1), np.float32)], 1)
    return (pts.astype(np.float64) @ t.T)


def boxes_to_image_space(dets, center, scale, out_w, out_h):
    """Both corners of every box [N,6] through the inverse of the network-input affine (what ctdet_post_process,
    lib/utils/post_process.py:86-103, does corner by corner): one 2x3 matrix, one matrix product per corner set."""
    t = get_affine_transform(center, scale, 0, (out_w, out_h), inv=1)
    d = np.array(dets, dtype=np.float32, copy=True)
    corners = d[:, :4].reshape(-1, 2).astype(np.float32)
    homog = np.concatenate([corners, np.ones((corners.shape[0], 1), np.float32)], 1).astype(np.float64)
    d[:, :4] = (homog @ t.T).reshape(-1, 4)
    return d


def group_by_class(dets, num_classes):
    """[N,6] (x1,y1,x2,y2,score,class) -> {1-based class id: float32 [n,5]} in detection order."""
    cls = dets[:, 5].astype(np.int64)
    return {j + 1: np.ascontiguousarray(dets[cls == j, :5], dtype=np.float32).reshape(-1, 5) for j in range(num_classes)}


def ctdet_post_process(dets, c, s, h, w, num_classes):
    """Host restatement used by the tests (the detector itself transforms on the device): per image, boxes back to image
    coordinates and grouped per 1-based class."""
    return [group_by_class(boxes_to_image_space(dets[i], c[i], s[i], w, h), num_classes) for i in range(dets.shape[0])]


class _Stages:
    """Wall-clock accounting of run(): the keys test.py prints (test.py:76-79)."""
    KEYS = ("load", "pre", "net", "dec", "post", "merge")

    def __init__(self):
        self.t = dict.fromkeys(self.KEYS, 0.0)
        self.start = self.mark = time.time()

    def lap(self, key, now=None):
        now = time.time() if now is None else now
        self.t[key] += now - self.mark
        self.mark = now

    def result(self, results):
        out = {"results": results, "tot": time.time() - self.start}
        out.update(self.t)
        return out


class CtdetDetector:
    def __init__(self, opt):
        if opt.gpus[0] < 0 or not torch.cuda.is_available():
            raise RuntimeError("codenet_b200 has no CPU execution path: CtdetDetector needs a B200 (opt.gpus >= 0)")
        opt.device = torch.device("cuda", opt.gpus[0])
        self.opt = opt
        self.float_model = not opt.resume_quantize
        self.mean = np.array(opt.mean, dtype=np.float32).reshape(1, 1, 3)
        self.std = np.array(opt.std, dtype=np.float32).reshape(1, 1, 3)
        self.max_per_image = 100
        self.num_classes = opt.num_classes
        self.scales = opt.test_scales
        if self.float_model:
            # the unquantised model (test.py without --resume-quantize): its state dict in the reference's key space feeds
            # engine_f32.EngineF32 (fp32 NCHW kernels; opt.f32_gemm = "fp32" | "tf32x3" selects the 1x1-conv arithmetic)
            sd = getattr(opt, "state_dict", None)
            if getattr(opt, "load_model", ""):
                ck = torch.load(opt.load_model, map_location="cpu")
                sd = ck["state_dict"] if "state_dict" in ck else ck
            if sd is None:
                raise RuntimeError("float CtdetDetector needs opt.load_model or opt.state_dict")
            self._f32_state = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
            self._f32_engines = {}
            self.model = None
            return
        self.model = PoseShuffleNetV2(opt.heads, opt.head_conv, w2=opt.w2, maxpool=opt.maxpool)
        quantize_shufflenetv2_dcn(self.model, quant_conv=opt.w_bit, quant_bn=None, quant_act=opt.a_bit,
                                  wt_quant_mode='symmetric', act_quant_mode='asymmetric', wt_per_channel=True,
                                  wt_percentile=opt.wt_percentile, act_percentile=opt.act_percentile,
                                  deform_backbone=False, w2=opt.w2, maxpool=opt.maxpool)
        if getattr(opt, "load_model", ""):
            self.load_model(opt.load_model)
        elif getattr(opt, "state_dict", None) is not None:
            self.load_state_dict(opt.state_dict)
        self.model.offset_mode = getattr(opt, "offset_mode", "bilinear")
        self.model.eval()
        freeze_ranges(self.model)
        self.mean = np.array(opt.mean, dtype=np.float32).reshape(1, 1, 3)
        self.std = np.array(opt.std, dtype=np.float32).reshape(1, 1, 3)
        self.max_per_image = 100
        self.num_classes = opt.num_classes
        self.scales = opt.test_scales
        self.opt = opt
        self.pause = True

    # -- checkpoint ingestion (lib/models/model.py:35-88: strips 'module.', tolerates the optimizer entry) ------------
    def load_state_dict(self, sd):
        sd = {(k[7:] if k.startswith("module.") else k): (v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v)))
              for k, v in sd.items()}
        own = self.model.state_dict()
        for k, v in sd.items():
            if k in own and tuple(own[k].shape) != tuple(v.shape):
                v = v.reshape(own[k].shape) if v.numel() == own[k].numel() else own[k]
                sd[k] = v
        missing, unexpected = self.model.load_state_dict(sd, strict=False)
        # the stage-shared QuantAct appears under every unit's name (all aliases of ONE module): a key is only missing
        # if none of its aliases was provided
        live = self.model.state_dict(keep_vars=True)
        given = {id(live[k]) for k in sd if k in live}
        missing = [k for k in missing if not k.endswith("num_batches_tracked") and id(live[k]) not in given]
        if missing:
            raise RuntimeError("checkpoint lacks %d tensors of the quantised CoDeNet graph, e.g. %s" % (len(missing), missing[:4]))
        return unexpected

    def load_model(self, path):
        ck = torch.load(path, map_location="cpu")
        return self.load_state_dict(ck["state_dict"] if "state_dict" in ck else ck)

    # -- geometry of the network input -------------------------------------------------------------------------------------
    def input_geometry(self, height, width, scale):
        """(input height, input width, centre, scale) of base_detector.py:48-60: fixed resolution -> the longer side is
        mapped onto the input width; otherwise the scaled image is padded up to the next multiple of pad + 1."""
        sh, sw = int(height * scale), int(width * scale)
        if self.opt.fix_res:
            return (sh, sw), (self.opt.input_h, self.opt.input_w), np.array([sw / 2., sh / 2.], np.float32), float(max(height, width))
        ih, iw = (sh | self.opt.pad) + 1, (sw | self.opt.pad) + 1
        return (sh, sw), (ih, iw), np.array([sw // 2, sh // 2], np.float32), np.array([iw, ih], np.float32)

    def _meta(self, c, s, ih, iw):
        r = self.opt.down_ratio
        return {'c': c, 's': s, 'out_height': ih // r, 'out_width': iw // r}

    def pre_process(self, image, scale, meta=None):
        """uint8 HWC image -> normalised fp32 [1|2,3,H,W] tensor + meta, cv2 resize / warpAffine as in the reference."""
        import cv2
        (sh, sw), (ih, iw), c, s = self.input_geometry(image.shape[0], image.shape[1], scale)
        warped = cv2.warpAffine(cv2.resize(image, (sw, sh)), get_affine_transform(c, s, 0, [iw, ih]), (iw, ih), flags=cv2.INTER_LINEAR)
        chw = ((warped / 255. - self.mean) / self.std).astype(np.float32).transpose(2, 0, 1)[None]
        if self.opt.flip_test:
            chw = np.concatenate([chw, chw[..., ::-1]], 0)
        return torch.from_numpy(np.ascontiguousarray(chw)), self._meta(c, s, ih, iw)

    def pre_process_device(self, image, scale=1, meta=None):
        """pre_process for a raw uint8 HWC image at test scale 1 ON THE DEVICE: the frame is uploaded as it is and
        cdn_warp_affine_u8 (bit-identical to the cv2.warpAffine call of base_detector.py:61-65; at scale 1 the cv2.resize in
        front of it is the identity) writes the network-sized uint8 image, mirrored copy included for --flip_test.  The
        normalisation then happens inside the stem kernel.  Returns (uint8 CUDA tensor [1|2, H, W, 3], meta)."""
        import ctypes as C
        from .. import _lib
        if scale != 1 or image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
            raise ValueError("pre_process_device takes a uint8 HWC image at test scale 1")
        (sh, sw), (ih, iw), c, s = self.input_geometry(image.shape[0], image.shape[1], 1)
        M = np.ascontiguousarray(get_affine_transform(c, s, 0, [iw, ih]), dtype=np.float64)
        src = torch.from_numpy(np.ascontiguousarray(image)).to(self.opt.device)
        n = 2 if self.opt.flip_test else 1
        dst = torch.empty((n, ih, iw, 3), dtype=torch.uint8, device=self.opt.device)
        with torch.cuda.device(self.opt.device):
            _lib.check(_lib.load().cdn_warp_affine_u8(C.c_void_p(src.data_ptr()), sh, sw, C.c_void_p(M.ctypes.data), C.c_void_p(dst.data_ptr()),
                                                      ih, iw, n - 1, C.c_void_p(torch.cuda.current_stream(self.opt.device).cuda_stream)))
        return dst, self._meta(c, s, ih, iw)

    def _is_identity_input(self, image, scale):
        """True when resize + warpAffine leave the image untouched: uint8, already input-sized, landscape or square (for a
        portrait image the longer side is the height, so the affine scales by input_w / height and pads)."""
        return (scale == 1 and self.opt.fix_res and not self.opt.flip_test and image.dtype == np.uint8 and
                image.shape[:2] == (self.opt.input_h, self.opt.input_w) and image.shape[1] >= image.shape[0])

    def _engine_for(self, H, W, batch, device):
        eng = self.model.compile_engine(H, W, max(batch, getattr(self.opt, "max_batch", 1)), device=device, K=self.opt.K)
        if not getattr(eng, "_norm", False):
            eng.set_normalization(self.mean.reshape(3), self.std.reshape(3))
        return eng

    def process_u8(self, images_u8, return_time=False):
        """Fast path of run(): uint8 [B,H,W,3] images already at the network's input size (then the reference's
        resize + warpAffine of pre_process are the identity and only the normalisation is left, which the stem kernel
        applies).  CUDA tensor in, same outputs as process()."""
        if self.float_model:
            return self._process_float(images_u8, return_time)
        B, H, W, _ = images_u8.shape
        dev = images_u8.device.index if images_u8.device.index is not None else torch.cuda.current_device()
        out = self._engine_for(H, W, B, dev).run(images_u8.contiguous(), maps=True, dets=True)
        torch.cuda.synchronize(images_u8.device)
        forward_time = time.time()
        output = {h: out[h] for h in self.opt.heads}
        return (output, out["dets"], forward_time) if return_time else (output, out["dets"])

    def _process_float(self, images, return_time):
        """process() for the unquantised model: EngineF32 forward (logits), hm.sigmoid_() as ctdet.py:32, decode on the device."""
        from ..arch import NetConfig
        from ..engine_f32 import EngineF32
        if images.dtype == torch.uint8:                              # pre_process_device's layout: normalise as base_detector.py:66-68
            mean = torch.as_tensor(self.mean.reshape(3), device=images.device)
            std = torch.as_tensor(self.std.reshape(3), device=images.device)
            images = ((images.float() / 255.0 - mean) / std).permute(0, 3, 1, 2)
        x = images.contiguous().float()
        B = x.shape[0]
        dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
        eng = self._f32_engines.get(dev)
        if eng is None:
            cfg = NetConfig(num_classes=int(self.opt.heads["hm"]), w2=bool(self.opt.w2), maxpool=bool(self.opt.maxpool))
            eng = self._f32_engines[dev] = EngineF32(cfg, self._f32_state, device=dev, K=self.opt.K,
                                                     gemm=getattr(self.opt, "f32_gemm", "fp32"))
        if not self.opt.flip_test and self.opt.reg_offset:
            dets, _, v = eng.detect(x)
            hm, wh, reg = v["hm"].sigmoid(), v["wh"], v.get("reg")
        else:
            v = eng.forward(x)
            hm, wh, reg = v["hm"].sigmoid(), v["wh"].contiguous(), (v.get("reg") if self.opt.reg_offset else None)
            if self.opt.flip_test:
                hm = (hm[0::2] + torch.flip(hm[1::2], [3])) / 2
                wh = (wh[0::2] + torch.flip(wh[1::2], [3])) / 2
                reg = reg[0::2] if reg is not None else None
            dets = ctdet_decode(hm, wh, reg=reg, cat_spec_wh=self.opt.cat_spec_wh, K=self.opt.K)
        output = {"hm": hm, "wh": wh}
        if reg is not None:
            output["reg"] = reg
        torch.cuda.synchronize(x.device)
        forward_time = time.time()
        return (output, dets, forward_time) if return_time else (output, dets)

    def process(self, images, return_time=False):
        """images: CUDA fp32 [B,3,H,W] (the reference's tensor) or uint8 [B,H,W,3] (pre_process_device).
        Returns (output {'hm' (post-sigmoid), 'wh', 'reg'}, dets [B|1,K,6][, time])."""
        if not images.is_cuda:
            raise RuntimeError("codenet_b200 has no CPU execution path: process() needs CUDA images")
        if self.float_model:
            return self._process_float(images, return_time)
        if images.dtype == torch.uint8:
            B, H, W, _ = images.shape
            x = images.contiguous()
        else:
            B, _, H, W = images.shape
            x = images.contiguous().float()
        dev = images.device.index if images.device.index is not None else torch.cuda.current_device()
        eng = self._engine_for(H, W, B, dev)
        if not self.opt.flip_test:
            # forward + sigmoid + decode in one graph replay; `hm` comes back post-sigmoid like output['hm'].sigmoid_()
            out = eng.run(x, maps=True, dets=True)
            torch.cuda.synchronize(images.device)
            forward_time = time.time()
            output = {h: out[h] for h in self.opt.heads}
            dets = out["dets"] if self.opt.reg_offset else ctdet_decode(out["hm"], out["wh"], reg=None, K=self.opt.K)
        else:
            # image + mirrored image: average hm and wh with the mirror flipped back (ctdet.py:35-38), on the device
            import ctypes as C
            from .. import _lib
            out = eng.run(x, maps=True, dets=False)
            output = {h: out[h] for h in self.opt.heads}
            pairs = B // 2
            hm = torch.empty((pairs,) + tuple(out["hm"].shape[1:]), dtype=torch.float32, device=images.device)
            wh = torch.empty((pairs,) + tuple(out["wh"].shape[1:]), dtype=torch.float32, device=images.device)
            with torch.cuda.device(images.device):
                _lib.check(_lib.load().cdn_ctdet_flip_merge(
                    C.c_void_p(out["hm"].data_ptr()), C.c_void_p(out["wh"].data_ptr()), pairs, hm.shape[1], hm.shape[2], hm.shape[3],
                    C.c_void_p(hm.data_ptr()), C.c_void_p(wh.data_ptr()),
                    C.c_void_p(torch.cuda.current_stream(images.device).cuda_stream)))
            reg = out["reg"][0::2].contiguous() if self.opt.reg_offset else None     # the unflipped image of every pair
            torch.cuda.synchronize(images.device)
            forward_time = time.time()
            dets = ctdet_decode(hm, wh, reg=reg, cat_spec_wh=self.opt.cat_spec_wh, K=self.opt.K)
        return (output, dets, forward_time) if return_time else (output, dets)

    def post_process(self, dets, meta, scale=1):
        """dets [1|B,K,6] in output-map coordinates -> {class: [n,5]} in image coordinates.  The affine runs on the device
        (cdn_ctdet_post_affine); the grouping per class is a view operation on the K rows that come back."""
        import ctypes as C
        from .. import _lib
        d = torch.as_tensor(dets, dtype=torch.float32, device=self.opt.device).detach().reshape(1, -1, 6).contiguous().clone()
        t = np.ascontiguousarray(get_affine_transform(meta['c'], meta['s'], 0, (meta['out_width'], meta['out_height']), inv=1),
                                 dtype=np.float64).reshape(1, 6)
        g = torch.empty_like(d)
        counts = torch.empty((1, self.num_classes), dtype=torch.int32, device=d.device)
        with torch.cuda.device(d.device):
            st = C.c_void_p(torch.cuda.current_stream(d.device).cuda_stream)
            L = _lib.load()
            _lib.check(L.cdn_ctdet_post_affine(C.c_void_p(d.data_ptr()), 1, d.shape[1], C.c_void_p(t.ctypes.data), st))
            _lib.check(L.cdn_ctdet_group_by_class(C.c_void_p(d.data_ptr()), 1, d.shape[1], self.num_classes, C.c_void_p(g.data_ptr()),
                                                  C.c_void_p(counts.data_ptr()), st))
        rows, cnt = g[0].cpu().numpy(), counts[0].cpu().numpy()
        ends = np.cumsum(cnt)
        out = {j + 1: np.ascontiguousarray(rows[ends[j] - cnt[j]:ends[j], :5], dtype=np.float32).reshape(-1, 5)
               for j in range(self.num_classes)}
        if scale != 1:
            for v in out.values():
                v[:, :4] /= scale
        return out

    def merge_outputs(self, detections):
        """Per class: all scales stacked; soft-NMS when several scales were run or --nms is set; then the max_per_image best
        scores overall survive (lib/detectors/ctdet.py:59-74)."""
        classes = range(1, self.num_classes + 1)
        merged = {j: np.ascontiguousarray(np.vstack([d[j] for d in detections]), dtype=np.float32) for j in classes}
        if len(self.scales) > 1 or self.opt.nms:
            for j in classes:
                soft_nms(merged[j], Nt=0.5, method=2)
        total = sum(len(merged[j]) for j in classes)
        if total > self.max_per_image:
            cut = np.partition(np.concatenate([merged[j][:, 4] for j in classes]), total - self.max_per_image)[total - self.max_per_image]
            merged = {j: v[v[:, 4] >= cut] for j, v in merged.items()}
        return merged

    def _image_of(self, src):
        """ndarray | path | pre-processed dict of the prefetching loader (test.py:62-72) -> (image, pre-processed dict or None)"""
        if isinstance(src, np.ndarray):
            return src, None
        if isinstance(src, str):
            import cv2
            return cv2.imread(src), None
        return src['image'][0].numpy(), src

    def run(self, image_or_path_or_tensor, meta=None):
        clock = _Stages()
        image, prepared = self._image_of(image_or_path_or_tensor)
        clock.lap("load")
        per_scale = []
        for scale in self.scales:
            if prepared is not None:
                images = prepared['images'][scale][0]
                meta = {k: (v.numpy()[0] if torch.is_tensor(v) else v) for k, v in prepared['meta'][scale].items()}
                fn = self.process
            elif self._is_identity_input(image, scale):
                # ship the uint8 image as it is; the stem kernel normalises through the 3 x 256 table
                h, w = image.shape[:2]
                images = torch.from_numpy(np.ascontiguousarray(image[None]))
                meta = self._meta(np.array([w / 2., h / 2.], np.float32), float(max(h, w)), h, w)
                fn = self.process_u8
            elif scale == 1 and image.dtype == np.uint8 and image.ndim == 3 and getattr(self.opt, "device_pre_process", True):
                images, meta = self.pre_process_device(image, scale, meta)           # warpAffine (+ mirror) on the device
                fn = self.process
            else:
                images, meta = self.pre_process(image, scale, meta)
                fn = self.process
            images = images.to(self.opt.device)
            torch.cuda.synchronize()
            clock.lap("pre")
            output, dets, t_forward = fn(images, return_time=True)
            torch.cuda.synchronize()
            clock.lap("net", t_forward)
            clock.lap("dec")
            per_scale.append(self.post_process(dets, meta, scale))
            clock.lap("post")
        results = self.merge_outputs(per_scale)
        clock.lap("merge")
        return clock.result(results)
