"""Value types of the module-level (B2) execution path: what flows between the Quant* modules of codenet_b200.compat when a
quantised network -- or a single block of it -- is run module by module instead of as one compiled engine.

The reference's modules pass fp32 NCHW tensors holding fake-quantised values (portable_quantizer/quant_modules.py).  Here

  QTensor      an activation on the 8-bit grid of ONE QuantAct: int8 NHWC on the device + (scale, zero point), real value
               (q + zero) / scale; `dequantize()` gives the reference's fp32 NCHW view on request.  `up` marks a pending
               nearest x2 upsample (shufflenetv2_dcn.py:303): it is virtual, the consumer reads (y >> 1, x >> 1).
  PendingConv  a conv (+ folded BatchNorm) whose output grid is not known yet: the reference's conv -> ReLU -> QuantAct chain is
               ONE kernel here (requantisation lives in the conv's epilogue), so QuantBnConv2d.forward returns this record,
               nn.ReLU marks it, and the following QuantAct.forward launches the fused kernel.

Both implement __torch_function__, so the unmodified nn.ReLU / nn.Upsample / nn.MaxPool2d members of the reference's module
tree (quantize_model.py:26-82) work on them.  No arithmetic runs on the CPU and nothing here falls back to eager PyTorch math.
"""
import numpy as np
import torch
import torch.nn.functional as F_


class QTensor:
    def __init__(self, q, C, act, half=0, up=0):
        assert q.dtype == torch.int8 and q.is_cuda and q.dim() == 4
        self.q, self.C, self.act, self.half, self.up = q, int(C), (float(act[0]), float(act[1])), int(half), int(up)

    # -- geometry -----------------------------------------------------------------------------------------------------------
    @property
    def pitch(self):
        return self.q.shape[3]

    @property
    def shape(self):
        B, H, W, _ = self.q.shape
        return (B, self.C, H << self.up, W << self.up)

    @property
    def device(self):
        return self.q.device

    def phys(self, c):
        """byte position of logical channel(s) c inside a pixel (HALF layout of the ShuffleNetV2 stages, plan.TensorSpec)"""
        c = np.asarray(c)
        if not self.half:
            return c
        h = self.C // 2
        return np.where(c < h, c, self.half + c - h)

    def spec(self, plan, name=""):
        return plan.add_tensor(self.q.shape[1], self.q.shape[2], self.C, self.pitch, self.half, self.act, name)

    def same_grid(self, act):
        return abs(self.act[0] - act[0]) <= 1e-12 * abs(act[0]) and self.act[1] == act[1]

    # -- the reference's view --------------------------------------------------------------------------------------------------
    def int_values(self):
        """logical int8 grid, NCHW (at the logical resolution)"""
        idx = torch.as_tensor(self.phys(np.arange(self.C)), device=self.q.device, dtype=torch.long)
        v = self.q.index_select(3, idx).permute(0, 3, 1, 2)
        for _ in range(self.up):
            v = v.repeat_interleave(2, 2).repeat_interleave(2, 3)
        return v.contiguous()

    def dequantize(self):
        """fp32 NCHW tensor of the real values (q + zero) / scale: what the reference's module would have returned"""
        s, z = self.act
        return ((self.int_values().double() + z) / s).float()

    def __repr__(self):
        return "QTensor(shape=%s, scale=%.6g, zero=%g, pitch=%d, half=%d, up=%d)" % (self.shape, self.act[0], self.act[1], self.pitch,
                                                                                     self.half, self.up)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", str(func))
        x = args[0]
        if name == "interpolate":
            sf = kwargs.get("scale_factor", args[2] if len(args) > 2 else None)
            mode = kwargs.get("mode", args[3] if len(args) > 3 else "nearest")
            if mode != "nearest" or sf not in (2, 2.0, (2, 2), (2.0, 2.0)) or x.up:
                raise NotImplementedError("QTensor: only one pending nearest x2 upsample is supported (shufflenetv2_dcn.py:303)")
            return QTensor(x.q, x.C, x.act, x.half, x.up + 1)
        if "max_pool2d" in name:
            k = kwargs.get("kernel_size", args[1] if len(args) > 1 else None)
            st = kwargs.get("stride", args[2] if len(args) > 2 else None)
            pd = kwargs.get("padding", args[3] if len(args) > 3 else 0)
            one = lambda v: v if isinstance(v, int) else (v[0] if v[0] == v[1] else None)
            if (one(k), one(st), one(pd)) != (3, 2, 1) or x.up:
                raise NotImplementedError("QTensor: only MaxPool2d(3, 2, 1) is on the path (quantize_model.py:31-35)")
            from .. import ops
            return QTensor(ops.maxpool3s2(x.q), x.C, x.act, x.half, 0)     # monotone: commutes with the quantiser
        if name in ("relu", "relu_"):
            raise NotImplementedError("QTensor: ReLU on an already quantised activation is not part of the graph; it follows a conv")
        raise NotImplementedError("QTensor does not implement torch.%s: call .dequantize() for an fp32 view" % name)


class PendingConv:
    """conv (+BN) of `module` applied to `x`, waiting for the ReLU / QuantAct that decide its output grid."""

    def __init__(self, module, x, relu=False):
        self.module, self.x, self.relu = module, x, bool(relu)

    @property
    def shape(self):
        return self.module._out_shape(self.x)

    def dequantize(self):
        """fp32 NCHW result of the conv itself (the reference's return value); 1x1 convs only"""
        y = self.module._float_out(self.x)
        return torch.relu_(y) if self.relu else y

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        name = getattr(func, "__name__", str(func))
        if name in ("relu", "relu_"):
            return PendingConv(args[0].module, args[0].x, True)
        raise NotImplementedError("PendingConv: torch.%s before the QuantAct that closes the conv; call .dequantize() for fp32" % name)
