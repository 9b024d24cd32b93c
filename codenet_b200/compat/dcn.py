"""Deformable-convolution layer classes of the reference, backed by libcodenet_b200.

Mirrors lib/models/external/functions/dcn_deform_conv.py:14-110 (DeformConvFunction, `deform_conv`) and
lib/models/external/modules/dcn_deform_conv.py:12-58 (DeformConv), :285-330 (DeformConvWithOffsetScaleBoundPositive).
The op-level call replaces `_ext.dcn.dcn_deform_conv_cuda.deform_conv_forward_cuda`
(lib/models/external/src/dcn_deform_conv_cuda.cpp:151-258) with cdn_deform_conv_forward_f32: gather and MAC fused,
no im2col buffer, no per-group addmm, launched on PyTorch's current stream.
"""
import ctypes as C
import math

import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from .. import _lib


def _output_size(input, weight, padding, dilation, stride):
    """functions/dcn_deform_conv.py:96-110."""
    channels = weight.size(0)
    output_size = (input.size(0), channels)
    for d in range(input.dim() - 2):
        in_size = input.size(d + 2)
        pad = padding[d]
        kernel = dilation[d] * (weight.size(d + 2) - 1) + 1
        output_size += ((in_size + (2 * pad) - kernel) // stride[d] + 1,)
    if not all(map(lambda s: s > 0, output_size)):
        raise ValueError("convolution input is too small (output would be {})".format("x".join(map(str, output_size))))
    return output_size


class DeformConvFunction:
    """Forward-only stand-in for the reference's autograd Function (inference path; backward is out of scope)."""

    @staticmethod
    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                im2col_step=64):
        if input is not None and input.dim() != 4:
            raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
        stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)
        if not input.is_cuda:
            raise NotImplementedError          # the reference refuses CPU tensors too (functions/dcn_deform_conv.py:43-45)
        if input.dtype != torch.float32:
            raise TypeError("codenet_b200 deform_conv: float32 only, got %s" % input.dtype)
        cur_im2col_step = min(im2col_step, input.shape[0])
        assert (input.shape[0] % cur_im2col_step) == 0, 'im2col step must divide batchsize'
        x, off, w = input.contiguous(), offset.contiguous().float(), weight.contiguous().float()
        output = x.new_empty(_output_size(x, w, padding, dilation, stride))
        if offset.shape[0] != x.shape[0] or offset.shape[2:] != output.shape[2:] or \
                offset.shape[1] != 2 * w.size(2) * w.size(3) * deformable_groups:
            raise RuntimeError("invalid offset shape {} for output {}".format(tuple(offset.shape), tuple(output.shape)))
        L = _lib.load()
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with torch.cuda.device(x.device):
            _lib.check(L.cdn_deform_conv_forward_f32(
                C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(off.data_ptr()),
                C.c_void_p(output.data_ptr()), x.shape[0], x.shape[1], x.shape[2], x.shape[3], w.shape[0],
                w.size(3), w.size(2), stride[1], stride[0], padding[1], padding[0], dilation[1], dilation[0],
                groups, deformable_groups, cur_im2col_step, stream))
        return output

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError("codenet_b200 is an inference engine: deform_conv backward is out of scope")

    @classmethod
    def apply(cls, *args, **kwargs):
        with torch.no_grad():
            return cls.forward(None, *args, **kwargs)


deform_conv = DeformConvFunction.apply


class DeformConv(nn.Module):
    """modules/dcn_deform_conv.py:12-58."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, dilation=1, groups=1,
                 deformable_groups=1, bias=False):
        super().__init__()
        assert not bias
        assert in_channels % groups == 0, 'in_channels {} cannot be divisible by groups {}'.format(in_channels, groups)
        assert out_channels % groups == 0, 'out_channels {} cannot be divisible by groups {}'.format(out_channels, groups)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = _pair(kernel_size), _pair(stride)
        self.padding, self.dilation = _pair(padding), _pair(dilation)
        self.groups, self.deformable_groups = groups, deformable_groups
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // self.groups, *self.kernel_size))
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, x, offset):
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)


ANCHOR_OFFSET = (-1, -1, -1, 0, -1, 1, 0, -1, 0, 0, 0, 1, 1, -1, 1, 0, 1, 1)


class DeformConvWithOffsetScaleBoundPositive(nn.Module):
    """The co-designed deformable module (modules/dcn_deform_conv.py:285-330): one scale scalar per output pixel,
    bounded to [-bound+1, bound], offsets = anchor * (s - 1), depthwise DeformConv, optional 1x1 channel conv.
    `groups`, `hidden_state` and `BN_MOMENTUM` are accepted and ignored exactly as in the reference."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, dilation=1, groups=1,
                 deformable_groups=1, bias=False, offset_bound=8, hidden_state=64, BN_MOMENTUM=0.1):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.offset_bound = offset_bound
        self.conv_scale = nn.Conv2d(in_channels, deformable_groups, kernel_size=1, stride=stride, padding=0, bias=True)
        self.conv_scale.weight.data.zero_()
        nn.init.constant_(self.conv_scale.bias, 1)
        self.conv_bound = nn.Hardtanh(min_val=-offset_bound + 1, max_val=offset_bound, inplace=True)
        self.conv = DeformConv(in_channels, in_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                               dilation=dilation, groups=in_channels, deformable_groups=deformable_groups, bias=bias)
        if in_channels != out_channels:
            self.conv_channel = nn.Conv2d(in_channels, out_channels, 1, 1, 0, bias=False)
            nn.init.kaiming_normal_(self.conv_channel.weight, nonlinearity='relu')
        self.anchor_offset = torch.FloatTensor(ANCHOR_OFFSET).unsqueeze(0).unsqueeze(2).unsqueeze(2)

    def forward(self, x):
        """fp32 module forward on CUDA tensors through two kernels of libcodenet_b200: the fused scale-conv + Hardtanh +
        bilinear gather + depthwise MAC (cdn_deform_dw_f32: no offset tensor, no im2col) and, when in != out, the 1x1
        conv_channel (cdn_pw_f32)."""
        if x.dim() != 4:
            raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(x.dim()))
        if not x.is_cuda:
            raise NotImplementedError          # as the reference's DeformConvFunction does for CPU tensors
        if x.dtype != torch.float32 or self.conv_scale.out_channels != 1 or self.conv.kernel_size != (3, 3) \
                or self.conv.padding != (1, 1) or self.conv.dilation != (1, 1):
            raise NotImplementedError("codenet_b200: the fused fp32 module covers CoDeNet's configuration "
                                      "(float32, 3x3, pad 1, dilation 1, deformable_groups 1)")
        L = _lib.load()
        x = x.contiguous()
        B, Cc, H, W = x.shape
        st = self.conv.stride[0]
        Ho, Wo = (H + 2 - 3) // st + 1, (W + 2 - 3) // st + 1
        y = x.new_empty((B, Cc, Ho, Wo))
        ws = self.conv_scale.weight.detach().reshape(-1).contiguous().float()
        wd = self.conv.weight.detach().reshape(Cc, 9).contiguous().float()
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with torch.cuda.device(x.device):
            _lib.check(L.cdn_deform_dw_f32(C.c_void_p(x.data_ptr()), C.c_void_p(ws.data_ptr()),
                                           C.c_float(float(self.conv_scale.bias.detach()[0])), int(self.offset_bound),
                                           C.c_void_p(wd.data_ptr()), C.c_void_p(y.data_ptr()), B, Cc, H, W, st, stream))
            if self.in_channels == self.out_channels:
                return y
            wc = self.conv_channel.weight.detach().reshape(self.out_channels, Cc).contiguous().float()
            out = x.new_empty((B, self.out_channels, Ho, Wo))
            _lib.check(L.cdn_pw_f32(C.c_void_p(y.data_ptr()), C.c_void_p(wc.data_ptr()), None, C.c_void_p(out.data_ptr()),
                                    B, Cc, self.out_channels, Ho * Wo, stream))
        return out
