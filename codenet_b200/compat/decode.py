"""ctdet decode mirror: lib/models/decode.py:474-505 (`ctdet_decode`, with `_nms` :10-16 and `_topk` :110-126 fused
into one pass on the GPU by cdn_ctdet_decode_prob).

`heat` is the POST-sigmoid heat map, exactly what the reference passes (lib/detectors/ctdet.py:32-41).  Peaks are
elements equal to the maximum of their 3x3 neighbourhood; the K best are returned in score order.  torch.topk leaves
the order of equal scores unspecified; this implementation fixes it to (score desc, class asc, index asc).
"""
import ctypes as C

import torch

from .. import _lib

_WS = {}


def ctdet_decode(heat, wh, reg=None, cat_spec_wh=False, K=100):
    if cat_spec_wh:
        raise NotImplementedError("cat_spec_wh=True is not used by CoDeNet (lib/opts.py) and not built")
    if not heat.is_cuda:
        raise RuntimeError("codenet_b200 has no CPU execution path: ctdet_decode needs CUDA tensors")
    batch, cat, height, width = heat.size()
    heat_c, wh_c = heat.contiguous().float(), wh.contiguous().float()
    reg_c = reg.contiguous().float() if reg is not None else None
    if wh_c.shape != (batch, 2, height, width) or (reg_c is not None and reg_c.shape != wh_c.shape):
        raise RuntimeError("ctdet_decode: wh / reg must be [B,2,H,W] matching heat {}".format(tuple(heat.shape)))
    dets = torch.empty((batch, K, 6), dtype=torch.float32, device=heat.device)
    L = _lib.load()
    need = int(L.cdn_ctdet_decode_ws_bytes(batch, cat, height, width))
    ws = _WS.get(heat.device)
    if ws is None or ws.numel() < need:              # candidate buffer kept between calls: the call only enqueues kernels
        ws = _WS[heat.device] = torch.empty(need, dtype=torch.uint8, device=heat.device)
    stream = C.c_void_p(torch.cuda.current_stream(heat.device).cuda_stream)
    with torch.cuda.device(heat.device):
        _lib.check(L.cdn_ctdet_decode_ws(C.c_void_p(heat_c.data_ptr()), 0, C.c_void_p(wh_c.data_ptr()), 0,
                                         C.c_void_p(reg_c.data_ptr()) if reg_c is not None else None, 0,
                                         batch, cat, height, width, K, 1, C.c_void_p(dets.data_ptr()), None,
                                         C.c_void_p(ws.data_ptr()), ws.numel(), stream))
    return dets
