"""Graph rewrite mirror: portable_quantizer/quantization_utils/quantize_model.py:7-82.

Rewrites a PoseShuffleNetV2 in place into the quantised module tree (same attribute names => same state-dict keys as
the reference, so `model_last.pth` checkpoints load): layer0 gets 8-bit weights (:28), every stage shares ONE output
QuantAct (:40), the three deformable blocks become QuantDeformConvWithOffsetScaleBoundPositive (:70-82).
"""
import torch.nn as nn

from .quant_modules import (QuantAct, QuantBnConv2d, QuantBaseNode, QuantDepthwiseNode,
                            QuantDeformConvWithOffsetScaleBoundPositive)


def quantize_shufflenetv2_dcn(model, quant_conv, quant_bn, quant_act, wt_quant_mode, act_quant_mode, wt_per_channel,
                              wt_percentile, act_percentile, deform_backbone, w2=False, maxpool=False):
    if deform_backbone:
        raise NotImplementedError("deform_backbone=True imports a module that does not exist in the reference "
                                  "(quant_modules.py:982, SURVEY.md F2); not built")
    if wt_quant_mode != "symmetric" or not wt_per_channel:
        raise NotImplementedError("the engine implements the reference's inference configuration: symmetric per-channel "
                                  "weights (lib/detectors/base_detector.py:29-34)")
    # wt_percentile changes the weight ranges at plan time (plan.weight_range); act_percentile only changes how QuantAct
    # collects its running range (quant_modules.py:203-219) and has no effect once the ranges are frozen
    model.wt_percentile = bool(wt_percentile)
    kw = dict(act_percentile=act_percentile, wt_quant_mode=wt_quant_mode, act_quant_mode=act_quant_mode,
              per_channel=wt_per_channel, weight_percentile=wt_percentile)

    def act():
        return QuantAct(quant_act, quant_mode="asymmetric", percentile=act_percentile)

    layer0 = model.layer0
    q0 = QuantBnConv2d(8, quant_mode=wt_quant_mode, per_channel=wt_per_channel, weight_percentile=wt_percentile)
    q0.set_param(layer0[0], layer0[1])
    tail = [layer0[2], act()] + ([layer0[3]] if maxpool else [])
    model.layer0 = nn.Sequential(q0, nn.Sequential(*tail))

    for n in range(1, 4):
        share_act = act()
        mods = []
        for node in getattr(model, 'layer%d' % n).children():
            qn = QuantBaseNode(quant_conv, quant_act, **kw)
            qn.set_param(node)
            qn.set_act(share_act)
            mods.append(qn)
        setattr(model, 'layer%d' % n, nn.Sequential(*mods))

    layer4 = model.layer4
    q4 = QuantBnConv2d(quant_conv, quant_mode=wt_quant_mode, per_channel=wt_per_channel, weight_percentile=wt_percentile)
    q4.set_param(layer4[0], layer4[1])
    model.layer4 = nn.Sequential(q4, nn.Sequential(layer4[2], act()))

    for head in model.heads:
        qh = QuantDepthwiseNode(quant_conv, quant_act, **kw)
        qh.set_param(getattr(model, head))
        setattr(model, head, qh)

    deform = model.deconv_layers
    mods = []
    for i in range(3):
        qd = QuantDeformConvWithOffsetScaleBoundPositive(quant_conv, quant_act, **kw)
        qd.set_param(deform[4 * i], deform[4 * i + 1])
        mods += [qd, nn.Sequential(deform[4 * i + 2], act()), deform[4 * i + 3]]
    model.deconv_layers = nn.Sequential(*mods)
    return model


def freeze_ranges(model):
    """Stop the running range statistics of every QuantAct: required before the graph is compiled."""
    n = 0
    for m in model.modules():
        if isinstance(m, QuantAct):
            m.running_stat = False
            n += 1
    return n
