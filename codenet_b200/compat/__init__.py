"""Host-side mirror of the reference's interface for the hot path (boundaries B1-B3).

Same class / function names, constructor arguments, state-dict keys and error behaviour as the reference, so that
checkpoints load unchanged and tests read like the reference's own usage -- but nothing here computes on the CPU:

  * dcn.py               deform_conv / DeformConvFunction / DeformConv / DeformConvWithOffsetScaleBoundPositive
                         (lib/models/external/{functions,modules}/dcn_deform_conv.py)
  * shufflenetv2_dcn.py  BaseNode / PoseShuffleNetV2 / get_shufflenetv2_dcn (lib/models/networks/shufflenetv2_dcn.py)
  * quant_modules.py     QuantAct / Quant_Conv2d / QuantBnConv2d / QuantLinear / QuantDeformConv2d /
                         QuantDeformConvWithOffsetScaleBoundPositive / QuantBaseNode / QuantDepthwiseNode
                         (portable_quantizer/quant_modules.py)
  * quantize_model.py    quantize_shufflenetv2_dcn (portable_quantizer/quantization_utils/quantize_model.py)
  * decode.py            ctdet_decode (lib/models/decode.py)
  * nms.py               soft_nms (lib/models/external/nms.pyx; host code in the reference too)
  * detector.py          CtdetDetector with run() / process() / pre_process() / post_process() / merge_outputs()
                         (lib/detectors/{base_detector,ctdet}.py)

The quantised modules are parameter containers: the reference evaluates them eagerly, one fake-quant tensor at a
time; here the whole quantised graph is compiled once into an int8 plan (codenet_b200.plan) and executed by
libcodenet_b200 on the GPU.  forward() of an individual quantised module runs its own kernels through the C ABI on QTensor
values (int8 NHWC + scale / zero point; qtensor.py, module_exec.py): boundary B2 at module granularity.
"""
from .dcn import deform_conv, DeformConvFunction, DeformConv, DeformConvWithOffsetScaleBoundPositive  # noqa: F401
from .shufflenetv2_dcn import BaseNode, PoseShuffleNetV2, get_shufflenetv2_dcn, channel_shuffle  # noqa: F401
from .quant_modules import (NotCompiledError, QuantAct, Quant_Conv2d, QuantBnConv2d, QuantLinear,  # noqa: F401
                            QuantDeformConv2d, QuantDeformConvWithOffsetScaleBoundPositive, QuantBaseNode,
                            QuantDepthwiseNode)
from .quantize_model import quantize_shufflenetv2_dcn, freeze_ranges  # noqa: F401
from .decode import ctdet_decode  # noqa: F401
from .nms import soft_nms  # noqa: F401
from .detector import CtdetDetector  # noqa: F401
from .qtensor import QTensor, PendingConv  # noqa: F401
