"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import-shim harness that runs the UNMODIFIED reference (Zhen-Dong/CoDeNet, mounted read-only at
/root/reference) on CPU, in fp32 or fp64, so that golden vectors can be generated from the reference's own
code.  It only works inside the build container (the GPU box has no /root/reference);
the vectors it produces are committed under tests/golden/ together with oracle/make_golden.py.

What is shimmed (no arithmetic lives in any of these):
  * pytorchcv{,.model_provider,.models,.models.shufflenetv2,.models.common}, thop  -- imported by
    lib/models/networks/shufflenetv2_dcn.py:12-13 and portable_quantizer/quant_modules.py:10-11, absent here.
  * _ext.dcn.dcn_deform_conv_cuda -- the reference's CUDA-only extension (functions/dcn_deform_conv.py:8,43-45
    refuses CPU tensors).  The name `deform_conv` captured by modules/dcn_deform_conv.py:9 and
    quant_modules.py:18 is rebound to torchvision.ops.deform_conv2d, a third-party CPU kernel with the same
    offset layout, range test and corner rule as dcn_deform_conv_cuda_kernel.cu:83-114,198-241.
"""
import os
import sys
import types

import torch
import torch.nn as nn

def _find_ref():
    """/root/reference in the build container; on the GPU box the unmodified copy that oracle/build_ref.build_py()
    installed into the git-ignored oracle/_ref/py (it travels with the repository snapshot)."""
    env = os.environ.get("CODENET_REFERENCE")
    if env:
        return env
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in ("/root/reference", os.path.join(here, "_ref", "py")):
        if os.path.isdir(os.path.join(cand, "lib", "models")):
            return cand
    return "/root/reference"


REF = _find_ref()


def available():
    return os.path.isdir(os.path.join(REF, "lib", "models"))


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's modules (net, dmod, qm, quantize, decode)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF)
    import torchvision.ops as tvops

    def fake(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    fake("pytorchcv")
    fake("pytorchcv.model_provider", get_model=lambda *a, **k: None)
    fake("pytorchcv.models")
    fake("pytorchcv.models.shufflenetv2", ShuffleUnit=_Stub)
    fake("pytorchcv.models.common", ChannelShuffle=_Stub)
    fake("thop", profile=lambda *a, **k: (0, 0))
    fake("_ext")
    fake("_ext.dcn", dcn_deform_conv_cuda=types.SimpleNamespace())
    for p in (REF, os.path.join(REF, "lib")):
        if p not in sys.path:
            sys.path.insert(0, p)

    from models.networks import shufflenetv2_dcn as net
    from models.external.modules import dcn_deform_conv as dmod
    import portable_quantizer.quant_modules as qm
    from portable_quantizer import quantize_shufflenetv2_dcn
    from models import decode as dec

    def cpu_deform_conv(x, offset, weight, stride=1, padding=0, dilation=1, groups=1,
                        deformable_groups=1, im2col_step=64):
        return tvops.deform_conv2d(x, offset, weight, None, stride, padding, dilation)

    dmod.deform_conv = cpu_deform_conv
    qm.deform_conv = cpu_deform_conv
    _loaded.update(net=net, dmod=dmod, qm=qm, quantize=quantize_shufflenetv2_dcn, decode=dec,
                   deform_conv=cpu_deform_conv)
    return types.SimpleNamespace(**_loaded)


class RoundOffsets(nn.Module):
    """Appended after the reference's offset-scale QuantAct to obtain the integer-offset ("FPGA") mode.

    The reference's only rounding variant is DeformConvWithOffsetRound (modules/dcn_deform_conv.py:225-237,
    `.round_()` = half-to-even); SURVEY.md F3 defines the integer mode of the W4A8 path as the same rounding
    applied to the bounded, quantised scalar s.
    """

    def forward(self, s):
        return torch.round(s)


def build_reference_model(raw_state, heads, w2=False, maxpool=False, dtype=torch.float64):
    """PoseShuffleNetV2 (shufflenetv2_dcn.py:189) built directly (the factory drops w2/maxpool),
    loaded with `raw_state` (pre-quantisation key space)."""
    R = load_reference()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.net.PoseShuffleNetV2(heads, head_conv=64, w2=w2, deform=False, maxpool=maxpool)
    missing, unexpected = m.load_state_dict(raw_state, strict=False)
    missing = [k for k in missing if not k.endswith("num_batches_tracked")]
    assert not missing and not unexpected, (missing, unexpected)
    return m.to(dtype)


def quantize_reference_model(m, w2=False, maxpool=False, w_bit=4, a_bit=8):
    """quantize_shufflenetv2_dcn exactly as base_detector.py:29-34 calls it."""
    R = load_reference()
    R.quantize(m, w_bit, None, a_bit, "symmetric", "asymmetric", True, False, False, False,
               w2=w2, maxpool=maxpool)
    return m


def freeze_ranges(m):
    R = load_reference()
    for mod in m.modules():
        if isinstance(mod, R.qm.QuantAct):
            mod.running_stat = False


def set_integer_offsets(m, enable=True):
    """Insert/remove RoundOffsets behind the three offset-scale QuantActs (quant_modules.py:651-653)."""
    R = load_reference()
    for mod in m.modules():
        if isinstance(mod, R.qm.QuantDeformConvWithOffsetScaleBoundPositive):
            seq = list(mod.quant_act.children())
            seq = [c for c in seq if not isinstance(c, RoundOffsets)]
            if enable:
                seq.append(RoundOffsets())
            mod.quant_act = nn.Sequential(*seq)
