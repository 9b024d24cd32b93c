"""Quantisation module mirror: portable_quantizer/quant_modules.py.

Same names, constructor signatures, `set_param` / `set_act` protocol and state-dict keys as the reference
(QuantAct :163-225, Quant_Conv2d :228-321, QuantBnConv2d :324-419, QuantDeformConv2d :422-517,
QuantDeformConvWithOffsetScaleBoundPositive :621-671, QuantBaseNode :809-907, QuantDepthwiseNode :1013-1071,
QuantLinear :23-160).  The reference evaluates these eagerly in fake-quant fp32 and re-quantises every weight on
every call.  Here they work at two levels:

  * whole network: the plan compiler (codenet_b200.plan) reads them ONCE -- BN folding, weight quantisation and activation
    scales in fp64 with the reference's formulae (quantization_utils/quant_utils.py:31-82, :170-223) -- and the graph runs as
    one compiled engine (PoseShuffleNetV2.forward);
  * module by module (boundary B2): every forward() below runs on the GPU through the stand-alone kernels of the C ABI
    (compat/module_exec.py), passing QTensor values (int8 NHWC + scale / zero point, compat/qtensor.py) instead of the
    reference's fake-quantised fp32 tensors; `.dequantize()` gives the reference's fp32 view.  Results are bit-identical to
    the compiled engine.  QuantAct ranges must be frozen.  There is no CPU path.
"""
import torch
import torch.nn as nn
from torch.nn import Module, Parameter


from .qtensor import QTensor, PendingConv


class NotCompiledError(RuntimeError):
    pass


def _need_q(x, who):
    if not isinstance(x, QTensor):
        raise TypeError("codenet_b200.%s.forward takes a QTensor (an activation on a QuantAct grid); got %s -- quantise real "
                        "values with a QuantAct first" % (who, type(x).__name__))
    return x


def _check_mode(quant_mode):
    if quant_mode not in ("symmetric", "asymmetric"):
        raise ValueError("unknown quant mode: {}".format(quant_mode))


class QuantAct(Module):
    """Activation quantiser with range buffers x_min / x_max of shape [1] (quant_modules.py:163-200).

    The reference keeps updating the range from every batch, even in eval mode (:203-219); the engine
    needs frozen ranges: load them from a checkpoint (they are buffers) or set them with `set_range`, then
    `running_stat = False` (see quantize_model.freeze_ranges)."""

    def __init__(self, activation_bit, momentum=0.99, full_precision_flag=False, running_stat=True,
                 quant_mode="symmetric", show_flag=False, percentile=False):
        super().__init__()
        self.activation_bit, self.momentum = activation_bit, momentum
        self.full_precision_flag, self.running_stat = full_precision_flag, running_stat
        self.quant_mode, self.show_flag, self.percentile = quant_mode, show_flag, percentile
        self.register_buffer('x_min', torch.zeros(1))
        self.register_buffer('x_max', torch.zeros(1))
        _check_mode(quant_mode)

    def set_range(self, x_min, x_max):
        self.x_min.fill_(float(x_min)); self.x_max.fill_(float(x_max))
        self.running_stat = False

    def __repr__(self):
        return "{0}(activation_bit={1}, full_precision_flag={2}, Act_min: {3:.2f}, Act_max: {4:.2f})".format(
            self.__class__.__name__, self.activation_bit, self.full_precision_flag, self.x_min.item(), self.x_max.item())

    def forward(self, x):
        """PendingConv -> the conv, its ReLU and this quantiser as ONE kernel; fp32 CUDA tensor (real values) -> quantised
        (cdn_quantize_f32_i8, quant_utils.py:31-39 with saturation); QTensor on this grid -> itself."""
        from . import module_exec as X
        act = X.act_of(self)
        if isinstance(x, PendingConv):
            return x.module._fused(x.x, act, x.relu)
        if isinstance(x, QTensor):
            if x.same_grid(act):
                return x
            raise NotImplementedError("QuantAct on a tensor that already lives on another grid is not part of the CoDeNet graph")
        if torch.is_tensor(x):
            if not x.is_cuda:
                raise RuntimeError("codenet_b200 has no CPU execution path: QuantAct needs a CUDA tensor")
            from .. import ops
            B, Cc, H, W = x.shape
            pitch = (Cc + 31) // 32 * 32
            return QTensor(ops.quantize(x, Cc, H, W, act[0], act[1], pitch), Cc, act)
        raise TypeError("QuantAct.forward: unsupported input %s" % type(x).__name__)


class Quant_Conv2d(Module):
    """Conv without BN (offset-scale conv, head output convs); set_param clones weight and bias (:263-276)."""

    def __init__(self, weight_bit, bias_bit=None, full_precision_flag=False, quant_mode="symmetric", per_channel=False,
                 weight_percentile=False):
        super().__init__()
        self.full_precision_flag, self.weight_bit, self.quant_mode = full_precision_flag, weight_bit, quant_mode
        self.momentum, self.per_channel, self.weight_percentile = 0.99, per_channel, weight_percentile
        self.bias_bit, self.quantize_bias = bias_bit, bias_bit is not None
        _check_mode(quant_mode)

    def set_param(self, conv):
        for a in ("in_channels", "out_channels", "kernel_size", "stride", "padding", "dilation", "groups"):
            setattr(self, a, getattr(conv, a))
        self.weight = Parameter(conv.weight.data.clone())
        try:
            self.bias = Parameter(conv.bias.data.clone())
        except AttributeError:
            self.bias = None

    # a Quant_Conv2d has no BatchNorm: `self` plays the conv's role for the shared helpers
    @property
    def _percentile(self):
        return bool(self.weight_percentile)

    def _out_shape(self, x):
        return (x.shape[0], self.out_channels, x.shape[2], x.shape[3])

    def _fused(self, x, act, relu):
        from . import module_exec as X
        return X.fused_conv(self, None, self.weight_bit, self._percentile, x, act, relu)

    def _float_out(self, x):
        from . import module_exec as X
        return X.float_conv(self, None, self.weight_bit, self._percentile, _need_q(x, "Quant_Conv2d"))

    def forward(self, x):
        """1x1 conv on a QTensor -> fp32 NCHW (the reference's use: head output convs and the offset-scale conv, neither is
        followed by a QuantAct that could close the kernel; quant_modules.py:278-321)"""
        return self._float_out(x)


class QuantBnConv2d(Module):
    """Conv + BatchNorm folded at compile time; set_param SHARES the original modules (:353-355)."""

    def __init__(self, weight_bit, bias_bit=None, full_precision_flag=False, running_stat=True, quant_mode="asymmetric",
                 per_channel=False, weight_percentile=False):
        super().__init__()
        self.weight_bit, self.full_precision_flag, self.running_stat = weight_bit, full_precision_flag, running_stat
        self.per_channel, self.weight_percentile = per_channel, weight_percentile
        self.bias_bit, self.quantize_bias = bias_bit, bias_bit is not None
        self.quant_mode = quant_mode
        _check_mode(quant_mode)

    def set_param(self, conv, bn):
        self.conv = conv
        self.bn = bn

    def _out_shape(self, x):
        st = self.conv.stride[0]
        return (x.shape[0], self.conv.out_channels, (x.shape[2] - 1) // st + 1, (x.shape[3] - 1) // st + 1)

    def _fused(self, x, act, relu):
        from . import module_exec as X
        return X.fused_conv(self.conv, self.bn, self.weight_bit, bool(self.weight_percentile), x, act, relu)

    def _float_out(self, x):
        from . import module_exec as X
        return X.float_conv(self.conv, self.bn, self.weight_bit, bool(self.weight_percentile), _need_q(x, "QuantBnConv2d"))

    def forward(self, x):
        """conv + folded BN (quant_modules.py:364-419).  The requantisation of the following QuantAct lives in the conv
        kernel's epilogue, so this returns a PendingConv that nn.ReLU marks and QuantAct.forward executes;
        `.dequantize()` on it gives the reference's fp32 return value (1x1 convs)."""
        if not (isinstance(x, QTensor) or (torch.is_tensor(x) and x.is_cuda)):
            raise TypeError("QuantBnConv2d.forward takes a QTensor (or the CUDA fp32 image for the stem conv)")
        return PendingConv(self, x)


class QuantDeformConv2d(Module):
    """Depthwise deformable conv weights, 4-bit per channel (:422-517); set_param clones the weight (:457-471)."""

    def __init__(self, weight_bit, bias_bit=None, full_precision_flag=False, quant_mode="symmetric", per_channel=False,
                 weight_percentile=False):
        super().__init__()
        self.full_precision_flag, self.weight_bit, self.quant_mode = full_precision_flag, weight_bit, quant_mode
        self.per_channel, self.weight_percentile = per_channel, weight_percentile
        self.bias_bit, self.quantize_bias = bias_bit, bias_bit is not None
        _check_mode(quant_mode)

    def set_param(self, conv):
        for a in ("in_channels", "out_channels", "kernel_size", "stride", "padding", "dilation", "groups",
                  "deformable_groups"):
            setattr(self, a, getattr(conv, a))
        self.weight = Parameter(conv.weight.data.clone())
        self.bias = None

    def forward(self, x, offset):
        """The general op of the reference (quant_modules.py:473-517 -> deform_conv): real-valued input (fp32 NCHW or a QTensor,
        dequantised), an explicit offset tensor [B,18,Ho,Wo], weights quantised per channel -> fp32 NCHW, through
        cdn_deform_conv_forward_f32.  The fused W4A8 layer is QuantDeformConvWithOffsetScaleBoundPositive.forward."""
        from .. import ops
        from ..plan import quant_weight
        xf = x.dequantize() if isinstance(x, QTensor) else x
        if not (torch.is_tensor(xf) and xf.is_cuda):
            raise RuntimeError("codenet_b200 has no CPU execution path: QuantDeformConv2d needs CUDA tensors")
        w = self.weight.detach().double().cpu().numpy()
        if self.full_precision_flag:
            wd = w
        else:
            if not self.per_channel:
                raise NotImplementedError("per-tensor deformable weights are not used by CoDeNet (base_detector.py:29-34)")
            wq, sigma = quant_weight(w, self.weight_bit, bool(self.weight_percentile))
            wd = wq / sigma.reshape(-1, 1, 1, 1)
        wt = torch.from_numpy(wd.astype("float32")).to(xf.device)
        return ops.deform_conv_f32(xf, offset, wt, self.stride, self.padding, self.dilation, self.groups, self.deformable_groups)


class _Compound(Module):
    def __init__(self, weight_bit, act_bit, full_precision_flag=False, bias_bit=None, act_percentile=False,
                 wt_quant_mode='symmetric', act_quant_mode='symmetric', per_channel=False, weight_percentile=False):
        super().__init__()
        self.act_bit, self.weight_bit, self.bias_bit = act_bit, weight_bit, bias_bit
        self.quantize_bias = bias_bit is not None
        self.wt_quant_mode, self.act_quant_mode = wt_quant_mode, act_quant_mode
        self.full_precision_flag, self.act_percentile = full_precision_flag, act_percentile
        self.per_channel, self.weight_percentile = per_channel, weight_percentile

    def _bnconv(self, conv, bn):
        m = QuantBnConv2d(self.weight_bit, quant_mode=self.wt_quant_mode, per_channel=self.per_channel,
                          weight_percentile=self.weight_percentile)
        m.set_param(conv, bn)
        return m

    def _act(self, mode=None):
        return QuantAct(self.act_bit, quant_mode=mode or "asymmetric", percentile=self.act_percentile)


class QuantDeformConvWithOffsetScaleBoundPositive(_Compound):
    """Quantised co-designed deformable block (:621-671): the unit the fused kernel cdn_deform_dw_w4a8 implements."""

    def set_param(self, deform_conv, bn):
        self.quant_conv_scale = Quant_Conv2d(self.weight_bit, quant_mode=self.wt_quant_mode, per_channel=self.per_channel,
                                             weight_percentile=self.weight_percentile)
        self.quant_conv_scale.set_param(deform_conv.conv_scale)
        self.quant_act = nn.Sequential(deform_conv.conv_bound, self._act())
        self.quant_deform_conv = QuantDeformConv2d(self.weight_bit, quant_mode=self.wt_quant_mode,
                                                   per_channel=self.per_channel, weight_percentile=self.weight_percentile)
        self.quant_deform_conv.set_param(deform_conv.conv)
        self.quant_identity_deform = self._act(self.act_quant_mode)
        self.anchor_offset = deform_conv.anchor_offset.clone()
        self.offset_bound = getattr(deform_conv, "offset_bound", int(deform_conv.conv_bound.max_val))
        self.quant_conv_channel_bn = self._bnconv(deform_conv.conv_channel, bn)

    def forward(self, x):
        """quant_modules.py:668-671: scale conv + Hardtanh + QuantAct(s) + offsets + depthwise deformable conv + QuantAct as the
        fused kernel (cdn_deform_dw_w4a8; `self.offset_mode` = "bilinear", the reference's arithmetic, or "round"), then the
        1x1 conv_channel + BN as a PendingConv for the caller's ReLU + QuantAct, exactly where the reference returns."""
        from . import module_exec as X
        y = X.deform_block(self, _need_q(x, "QuantDeformConvWithOffsetScaleBoundPositive"))
        return self.quant_conv_channel_bn(y)


class QuantBaseNode(_Compound):
    """Quantised ShuffleNetV2 unit (:809-907); the stage-shared output QuantAct arrives through set_act."""

    def set_param(self, base_node):
        self.stride = base_node.stride
        b2 = base_node.b2
        self.quant_convbn1 = self._bnconv(b2[0], b2[1])
        self.quant_act1 = self._act()
        assert type(b2[3]) == nn.Conv2d
        self.quant_convbn2 = self._bnconv(b2[3], b2[4])
        self.quant_act2 = self._act(self.act_quant_mode)
        self.quant_convbn3 = self._bnconv(b2[5], b2[6])
        if base_node.stride == 2:
            b1 = base_node.b1
            assert type(b1[0]) == nn.Conv2d
            self.quant_convbn4 = self._bnconv(b1[0], b1[1])
            self.quant_act4 = self._act(self.act_quant_mode)
            self.quant_convbn5 = self._bnconv(b1[2], b1[3])

    def set_act(self, share_quant_act):
        self.quant_act = share_quant_act

    def forward(self, x):
        from . import module_exec as X
        return X.base_node(self, _need_q(x, "QuantBaseNode"))


class QuantDepthwiseNode(_Compound):
    """Quantised depthwise-separable head (:1013-1071)."""

    def set_param(self, head_node):
        self.quant_convbn1 = self._bnconv(head_node[0], head_node[1])
        self.quant_act1 = nn.Sequential(head_node[2], self._act())
        assert type(head_node[3]) == nn.Conv2d
        self.quant_convbn2 = self._bnconv(head_node[3], head_node[4])
        self.quant_act3 = nn.Sequential(head_node[5], self._act())
        self.quant_conv = Quant_Conv2d(self.weight_bit, quant_mode=self.wt_quant_mode, per_channel=self.per_channel,
                                       weight_percentile=self.weight_percentile)
        self.quant_conv.set_param(head_node[6])

    def forward(self, x):
        from . import module_exec as X
        return X.head_node(self, _need_q(x, "QuantDepthwiseNode"))


class QuantLinear(nn.Linear):
    """Kept for API parity (:23-160); CoDeNet never instantiates it, so no kernel backs it."""

    def __init__(self, weight_bit, input_size, output_size, full_precision_flag=False, quant_mode="symmetric",
                 per_channel=False, show_flag=False, weight_percentile=False, save_path=None, threshold=None):
        super().__init__(input_size, output_size)
        self.full_precision_flag, self.weight_bit, self.quant_mode = full_precision_flag, weight_bit, quant_mode
        self.per_channel, self.show_flag, self.weight_percentile = per_channel, show_flag, weight_percentile
        self.save_path, self.threshold = save_path, threshold
        _check_mode(quant_mode)

    def forward(self, x):
        raise NotCompiledError("QuantLinear is not instantiated by any CoDeNet graph (quantize_model.py never creates one); "
                               "no kernel backs it")
