"""Detector entry points mirror: lib/detectors/base_detector.py:20-155 and lib/detectors/ctdet.py:26-74.

`CtdetDetector(opt)` builds the network, quantises it, loads the checkpoint (`opt.load_model`, reference key space),
freezes the activation ranges and runs forward + decode through the compiled int8 engine.  `run()` returns the same
dict as the reference (`results` keyed by 1-based class id and the stage timers test.py prints, test.py:76-79).
Pre- and post-processing are host code as in the reference (SURVEY.md 8(f) rows 1-2 move them to the device later).
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


def ctdet_post_process(dets, c, s, h, w, num_classes):
    """lib/utils/post_process.py:86-103: boxes back to image coordinates, 1-based class dict per image."""
    ret = []
    for i in range(dets.shape[0]):
        dets[i, :, :2] = transform_preds(dets[i, :, 0:2], c[i], s[i], (w, h))
        dets[i, :, 2:4] = transform_preds(dets[i, :, 2:4], c[i], s[i], (w, h))
        classes = dets[i, :, -1]
        top = {}
        for j in range(num_classes):
            inds = classes == j
            top[j + 1] = np.concatenate([dets[i, inds, :4].astype(np.float32), dets[i, inds, 4:5].astype(np.float32)],
                                        axis=1).tolist()
        ret.append(top)
    return ret


class CtdetDetector:
    def __init__(self, opt):
        if opt.gpus[0] < 0 or not torch.cuda.is_available():
            raise RuntimeError("codenet_b200 has no CPU execution path: CtdetDetector needs a B200 (opt.gpus >= 0)")
        opt.device = torch.device("cuda", opt.gpus[0])
        if not opt.resume_quantize:
            raise NotImplementedError("only the quantised (W4A8) path is built; pass resume_quantize=True")
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

    # -- reference API ---------------------------------------------------------------------------------------------------
    def pre_process(self, image, scale, meta=None):
        import cv2
        height, width = image.shape[0:2]
        new_height, new_width = int(height * scale), int(width * scale)
        if self.opt.fix_res:
            inp_height, inp_width = self.opt.input_h, self.opt.input_w
            c = np.array([new_width / 2., new_height / 2.], dtype=np.float32)
            s = max(height, width) * 1.0
        else:
            inp_height, inp_width = (new_height | self.opt.pad) + 1, (new_width | self.opt.pad) + 1
            c = np.array([new_width // 2, new_height // 2], dtype=np.float32)
            s = np.array([inp_width, inp_height], dtype=np.float32)
        trans_input = get_affine_transform(c, s, 0, [inp_width, inp_height])
        resized = cv2.resize(image, (new_width, new_height))
        inp = cv2.warpAffine(resized, trans_input, (inp_width, inp_height), flags=cv2.INTER_LINEAR)
        inp = ((inp / 255. - self.mean) / self.std).astype(np.float32)
        images = inp.transpose(2, 0, 1).reshape(1, 3, inp_height, inp_width)
        if self.opt.flip_test:
            images = np.concatenate((images, images[:, :, :, ::-1]), axis=0)
        images = torch.from_numpy(np.ascontiguousarray(images))
        meta = {'c': c, 's': s, 'out_height': inp_height // self.opt.down_ratio,
                'out_width': inp_width // self.opt.down_ratio}
        return images, meta

    def _engine_for(self, H, W, batch, device):
        eng = self.model.compile_engine(H, W, max(batch, getattr(self.opt, "max_batch", 1)), device=device, K=self.opt.K)
        if not getattr(eng, "_norm", False):
            eng.set_normalization(self.mean.reshape(3), self.std.reshape(3))
        return eng

    def process_u8(self, images_u8, return_time=False):
        """Fast path of run(): uint8 [B,H,W,3] images already at the network's input size (then the reference's
        resize + warpAffine of pre_process are the identity and only the normalisation is left, which the stem kernel
        applies).  CUDA tensor in, same outputs as process()."""
        B, H, W, _ = images_u8.shape
        dev = images_u8.device.index if images_u8.device.index is not None else torch.cuda.current_device()
        out = self._engine_for(H, W, B, dev).run(images_u8.contiguous(), maps=True, dets=True)
        torch.cuda.synchronize(images_u8.device)
        forward_time = time.time()
        output = {h: out[h] for h in self.opt.heads}
        return (output, out["dets"], forward_time) if return_time else (output, out["dets"])

    def process(self, images, return_time=False):
        """images: CUDA fp32 [B,3,H,W].  Returns (output {'hm' (post-sigmoid), 'wh', 'reg'}, dets [B|1,K,6][, time])."""
        if not images.is_cuda:
            raise RuntimeError("codenet_b200 has no CPU execution path: process() needs CUDA images")
        B, _, H, W = images.shape
        dev = images.device.index if images.device.index is not None else torch.cuda.current_device()
        eng = self._engine_for(H, W, B, dev)
        x = images.contiguous().float()
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
            reg = out["reg"][0:1] if self.opt.reg_offset else None
            torch.cuda.synchronize(images.device)
            forward_time = time.time()
            dets = ctdet_decode(hm, wh, reg=reg, cat_spec_wh=self.opt.cat_spec_wh, K=self.opt.K)
        return (output, dets, forward_time) if return_time else (output, dets)

    def post_process(self, dets, meta, scale=1):
        if torch.is_tensor(dets) and dets.is_cuda and dets.dtype == torch.float32:
            # coordinate transform on the device (cdn_ctdet_post_affine), only the per-class grouping stays on the host
            import ctypes as C
            from .. import _lib
            d = dets.detach().reshape(1, -1, dets.shape[2]).contiguous().clone()
            t = np.ascontiguousarray(get_affine_transform(meta['c'], meta['s'], 0, (meta['out_width'], meta['out_height']),
                                                          inv=1), dtype=np.float64).reshape(1, 6)
            with torch.cuda.device(d.device):
                _lib.check(_lib.load().cdn_ctdet_post_affine(C.c_void_p(d.data_ptr()), 1, d.shape[1], C.c_void_p(t.ctypes.data),
                                                             C.c_void_p(torch.cuda.current_stream(d.device).cuda_stream)))
            d = d.cpu().numpy()
            classes = d[0, :, -1]
            out = {}
            for j in range(self.num_classes):
                out[j + 1] = np.ascontiguousarray(d[0, classes == j, :5], dtype=np.float32).reshape(-1, 5)
                out[j + 1][:, :4] /= scale
            return out
        dets = dets.detach().cpu().numpy()
        dets = dets.reshape(1, -1, dets.shape[2])
        dets = ctdet_post_process(dets.copy(), [meta['c']], [meta['s']], meta['out_height'], meta['out_width'],
                                  self.opt.num_classes)
        for j in range(1, self.num_classes + 1):
            dets[0][j] = np.array(dets[0][j], dtype=np.float32).reshape(-1, 5)
            dets[0][j][:, :4] /= scale
        return dets[0]

    def merge_outputs(self, detections):
        """lib/detectors/ctdet.py:59-74: concatenate the scales per class, soft-NMS (gaussian, Nt = 0.5, in place, as the
        reference calls it) when testing at several scales or with --nms, keep the max_per_image best."""
        results = {j: np.ascontiguousarray(np.concatenate([d[j] for d in detections], axis=0).astype(np.float32))
                   for j in range(1, self.num_classes + 1)}
        if len(self.scales) > 1 or self.opt.nms:
            for j in range(1, self.num_classes + 1):
                soft_nms(results[j], Nt=0.5, method=2)
        scores = np.hstack([results[j][:, 4] for j in range(1, self.num_classes + 1)])
        if len(scores) > self.max_per_image:
            kth = len(scores) - self.max_per_image
            thresh = np.partition(scores, kth)[kth]
            for j in range(1, self.num_classes + 1):
                results[j] = results[j][results[j][:, 4] >= thresh]
        return results

    def run(self, image_or_path_or_tensor, meta=None):
        load_time = pre_time = net_time = dec_time = post_time = merge_time = tot_time = 0
        start_time = time.time()
        pre_processed = False
        if isinstance(image_or_path_or_tensor, np.ndarray):
            image = image_or_path_or_tensor
        elif isinstance(image_or_path_or_tensor, str):
            import cv2
            image = cv2.imread(image_or_path_or_tensor)
        else:
            image = image_or_path_or_tensor['image'][0].numpy()
            pre_processed_images = image_or_path_or_tensor
            pre_processed = True
        loaded_time = time.time()
        load_time += loaded_time - start_time
        detections = []
        for scale in self.scales:
            scale_start_time = time.time()
            fast = (not pre_processed and scale == 1 and not self.opt.flip_test and self.opt.fix_res and
                    image.dtype == np.uint8 and image.shape[:2] == (self.opt.input_h, self.opt.input_w))
            if fast:
                # identity resize/affine: ship the uint8 image, normalise inside the stem kernel
                h, w = image.shape[:2]
                meta = {'c': np.array([w / 2., h / 2.], dtype=np.float32), 's': max(h, w) * 1.0,
                        'out_height': h // self.opt.down_ratio, 'out_width': w // self.opt.down_ratio}
                images = torch.from_numpy(np.ascontiguousarray(image[None]))
            elif not pre_processed:
                images, meta = self.pre_process(image, scale, meta)
            else:
                images = pre_processed_images['images'][scale][0]
                meta = pre_processed_images['meta'][scale]
                meta = {k: (v.numpy()[0] if torch.is_tensor(v) else v) for k, v in meta.items()}
            images = images.to(self.opt.device)
            torch.cuda.synchronize()
            pre_process_time = time.time()
            pre_time += pre_process_time - scale_start_time
            if fast:
                output, dets, forward_time = self.process_u8(images, return_time=True)
            else:
                output, dets, forward_time = self.process(images, return_time=True)
            torch.cuda.synchronize()
            net_time += forward_time - pre_process_time
            decode_time = time.time()
            dec_time += decode_time - forward_time
            dets = self.post_process(dets, meta, scale)
            post_process_time = time.time()
            post_time += post_process_time - decode_time
            detections.append(dets)
        results = self.merge_outputs(detections)
        end_time = time.time()
        merge_time += end_time - post_process_time
        tot_time += end_time - start_time
        return {'results': results, 'tot': tot_time, 'load': load_time, 'pre': pre_time, 'net': net_time,
                'dec': dec_time, 'post': post_time, 'merge': merge_time}
