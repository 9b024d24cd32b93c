"""Network definition mirror: lib/models/networks/shufflenetv2_dcn.py (BaseNode :57-114, PoseShuffleNetV2 :189-330,
get_shufflenetv2_dcn :364-373) with identical module names, hence identical state-dict keys.

The float modules only hold parameters (checkpoint compatibility, BN folding input).  forward() of the top-level
network runs the compiled int8 engine once the model has been quantised and its ranges are frozen; there is no
float / CPU execution path.
"""
import torch
import torch.nn as nn

from .dcn import DeformConvWithOffsetScaleBoundPositive

BN_MOMENTUM = 0.1


def channel_shuffle(x, groups):
    """shufflenetv2_dcn.py:29-34 (index arithmetic only; the engine folds it into the 1x1 conv epilogue)."""
    b, c, h, w = x.shape
    return x.view(b, groups, c // groups, h, w).transpose(1, 2).contiguous().view(b, -1, h, w)


class BaseNode(nn.Module):
    def __init__(self, inp, oup, stride, batch_norm, conv_kernel):
        super().__init__()
        self.stride = stride
        h = oup // 2

        def branch2(cin):
            return nn.Sequential(
                nn.Conv2d(cin, h, 1, 1, 0, bias=False), batch_norm(h, momentum=BN_MOMENTUM), nn.ReLU(inplace=True),
                conv_kernel(h, h, 3, stride, 1, groups=h, bias=False), batch_norm(h, momentum=BN_MOMENTUM),
                nn.Conv2d(h, h, 1, 1, 0, bias=False), batch_norm(h, momentum=BN_MOMENTUM), nn.ReLU(inplace=True))

        if stride == 1:
            self.b2 = branch2(h)
        elif stride == 2:
            self.b1 = nn.Sequential(
                conv_kernel(inp, inp, 3, 2, 1, groups=inp, bias=False), batch_norm(inp, momentum=BN_MOMENTUM),
                nn.Conv2d(inp, h, 1, 1, 0, bias=False), batch_norm(h, momentum=BN_MOMENTUM), nn.ReLU(inplace=True))
            self.b2 = branch2(inp)

    def forward(self, x):
        raise RuntimeError("codenet_b200: BaseNode holds parameters only; run the network through its engine")


class PoseShuffleNetV2(nn.Module):
    def __init__(self, heads, head_conv, w2=None, deform=False, maxpool=False):
        super().__init__()
        if deform:
            raise NotImplementedError("deform_backbone=True is dead code in the reference (SURVEY.md F2) and not built")
        self.w2, self.deform_backbone, self.heads, self.maxpool = w2, deform, heads, maxpool
        self.deconv_with_bias = False
        self.channels = [24, 244, 488, 976, 2153] if w2 is True else [24, 116, 232, 464, 1024]
        ch = self.channels
        l0 = [nn.Conv2d(3, ch[0], 3, 2 if maxpool else 4, 1, bias=False), nn.BatchNorm2d(ch[0], momentum=BN_MOMENTUM),
              nn.ReLU(inplace=True)]
        if maxpool:
            l0.append(nn.MaxPool2d(kernel_size=3, stride=2, padding=1))
        self.layer0 = nn.Sequential(*l0)
        for idx, reps in enumerate([3, 7, 3]):
            layers = [BaseNode(ch[idx], ch[idx + 1], 2, nn.BatchNorm2d, nn.Conv2d)]
            layers += [BaseNode(ch[idx], ch[idx + 1], 1, nn.BatchNorm2d, nn.Conv2d) for _ in range(reps)]
            setattr(self, 'layer' + str(idx + 1), nn.Sequential(*layers))
        self.layer4 = nn.Sequential(nn.Conv2d(ch[3], ch[4], 1, 1, 0, bias=False),
                                    nn.BatchNorm2d(ch[4], momentum=BN_MOMENTUM), nn.ReLU(inplace=True))
        self.deconv_layers = self._make_deconv_layer(3, [256, 128, 64], [3, 3, 3])
        for head in self.heads:
            classes = self.heads[head]
            if head_conv > 0:
                fc = nn.Sequential(
                    nn.Conv2d(64, head_conv, 1, 1, 0, bias=False), nn.BatchNorm2d(head_conv, momentum=BN_MOMENTUM),
                    nn.ReLU(inplace=True),
                    nn.Conv2d(head_conv, head_conv, 3, 1, 1, groups=head_conv, bias=False),
                    nn.BatchNorm2d(head_conv, momentum=BN_MOMENTUM), nn.ReLU(inplace=True),
                    nn.Conv2d(head_conv, classes, kernel_size=1, stride=1, padding=0, bias=True))
                if 'hm' in head:
                    fc[-1].bias.data.fill_(-2.19)
                else:
                    for m in fc.modules():
                        if isinstance(m, nn.Conv2d):
                            nn.init.kaiming_normal_(m.weight.data, nonlinearity='relu')
                            if m.bias is not None:
                                nn.init.constant_(m.bias, 0)
            else:
                raise NotImplementedError("head_conv == 0 is not used by CoDeNet")
            self.__setattr__(head, fc)
        self._engine = None
        self._engine_key = None
        self.offset_mode = "bilinear"          # the reference's behaviour; "round" = integer offsets

    def _make_deconv_layer(self, num_layers, num_filters, num_kernels):
        assert num_layers == len(num_filters) == len(num_kernels)
        planes_in = [2153, 256, 128] if self.w2 is True else [1024, 256, 128]
        layers = []
        for i in range(num_layers):
            planes = num_filters[i]
            layers += [DeformConvWithOffsetScaleBoundPositive(planes_in[i], planes, 3, 1, 1, groups=planes, bias=False,
                                                              hidden_state=128, BN_MOMENTUM=BN_MOMENTUM),
                       nn.BatchNorm2d(planes, momentum=BN_MOMENTUM), nn.ReLU(inplace=True),
                       nn.Upsample(size=None, scale_factor=2, mode='nearest', align_corners=None)]
        return nn.Sequential(*layers)

    # -- execution: compiled engine -----------------------------------------------------------------------------------
    def compile_engine(self, in_h, in_w, max_batch, device=0, K=100, offset_mode=None):
        """Compile the (quantised, range-frozen) network into an int8 plan on `device`."""
        from ..engine import Engine
        from .quant_modules import QuantAct
        acts = [m for m in self.modules() if isinstance(m, QuantAct)]
        if not acts:
            raise RuntimeError("codenet_b200 executes the W4A8 graph: call quantize_shufflenetv2_dcn(model, ...) first "
                               "(the fp32 model path is not built, DESIGN.md section 9)")
        if any(a.running_stat for a in acts):
            raise RuntimeError("QuantAct ranges are still running statistics (the reference updates them on every "
                               "forward, even in eval mode -- SURVEY.md F4); load calibrated ranges and call "
                               "freeze_ranges(model) before inference")
        mode = offset_mode or self.offset_mode
        key = (in_h, in_w, max_batch, device, K, mode)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine.from_module(self, in_h, in_w, max_batch, offset_mode=mode, device=device, K=K)
            self._engine_key = key
        return self._engine

    def forward(self, x):
        """x: CUDA fp32 [B,3,H,W].  Returns [{'hm' (logits), 'wh', 'reg'}] like the reference (:314-330)."""
        if not x.is_cuda:
            raise RuntimeError("codenet_b200 has no CPU execution path: move the input to a B200")
        e = self._engine
        if e is None or self._engine_key[:2] != (x.shape[2], x.shape[3]) or x.shape[0] > self._engine_key[2]:
            dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
            e = self.compile_engine(x.shape[2], x.shape[3], max(int(x.shape[0]), 1), device=dev)
        out = e.run(x.contiguous().float(), maps=True, dets=False, raw_hm=True)
        return [{h: out[h] for h in self.heads}]

    def forward_modules(self, x):
        """The reference's forward (shufflenetv2_dcn.py:314-330) module by module: every Quant* member runs its own kernels
        through the C ABI on QTensor values (boundary B2; compat/module_exec.py).  Same bits as forward(), many more launches:
        meant for graphs that keep the reference's module structure, not for speed.  x: CUDA fp32 [B,3,H,W]."""
        if not x.is_cuda:
            raise RuntimeError("codenet_b200 has no CPU execution path: move the input to a B200")
        from .quant_modules import QuantDeformConvWithOffsetScaleBoundPositive
        for m in self.modules():
            if isinstance(m, QuantDeformConvWithOffsetScaleBoundPositive):
                m.offset_mode = self.offset_mode
        x = self.layer0(x.contiguous().float())
        x = self.layer1(x)
        x = self.layer2(x)
        x = self.layer3(x)
        x = self.layer4(x)
        x = self.deconv_layers(x)
        return [{head: getattr(self, head)(x) for head in self.heads}]

    def init_weights(self, num_layers):
        """The reference builds a renamed state dict from a downloaded pytorchcv model and never loads it
        (shufflenetv2_dcn.py:332-361): a no-op here (no network access either)."""
        return None


def get_shufflenetv2_dcn(num_layers, heads, head_conv):
    """shufflenetv2_dcn.py:364-373 without its `.cuda()`, thop profiling and download side effects."""
    model = PoseShuffleNetV2(heads, head_conv=head_conv, deform=False)
    model.init_weights(num_layers)
    return model
