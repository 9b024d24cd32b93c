"""Static description of the CoDeNet graph (ShuffleNetV2 + co-designed deformable up-path + ctdet heads).

This is the host-side mirror of the reference's network definition
(`lib/models/networks/shufflenetv2_dcn.py:189-330`, `BaseNode` :57-114) and of the graph rewrite done by
`portable_quantizer/quantization_utils/quantize_model.py:7-82`: it only enumerates shapes and state-dict key
names -- in both the raw (pre-quantisation) and the quantised key space -- so that checkpoints written by the
reference load unchanged.  No arithmetic lives here.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

STAGE_REPEATS = (3, 7, 3)                 # shufflenetv2_dcn.py:214
UP_FILTERS = (256, 128, 64)               # shufflenetv2_dcn.py:238-242
HEAD_CONV = 64


@dataclass(frozen=True)
class NetConfig:
    """The knobs of the reference that reach the hot path (lib/opts.py:230-248)."""
    num_classes: int = 20
    w2: bool = False                       # 2x width ("CoDeNet2x")
    maxpool: bool = False                  # stem stride 2 + MaxPool(3,2,1) instead of stride 4
    w_bit: int = 4
    a_bit: int = 8
    offset_bound: int = 8                  # Hardtanh(-bound+1, bound), dcn_deform_conv.py:304-305
    wt_percentile: bool = False            # --wt-percentile: weight ranges from the 0.1 / 99.9 percentiles (quant_modules.py:382-395)
    heads: Tuple[Tuple[str, int], ...] = ()

    def head_list(self) -> List[Tuple[str, int]]:
        return list(self.heads) if self.heads else [("hm", self.num_classes), ("wh", 2), ("reg", 2)]

    @property
    def channels(self) -> List[int]:       # shufflenetv2_dcn.py:199-202
        return [24, 244, 488, 976, 2153] if self.w2 else [24, 116, 232, 464, 1024]

    @property
    def up_planes(self) -> List[int]:      # shufflenetv2_dcn.py:292-295
        return [self.channels[4], 256, 128]


@dataclass
class ConvSpec:
    """One convolution of the graph with the state-dict prefixes it reads."""
    name: str                              # our own label
    kind: str                              # 'stem' | 'pw' | 'dw' | 'deform_dw' | 'scale' | 'head_out'
    cin: int
    cout: int
    k: int
    stride: int
    groups: int
    raw_conv: str                          # key prefix of the conv in the raw model
    raw_bn: str                            # key prefix of its BatchNorm ('' if none)
    q_conv: str                            # key prefix of the conv weight in the quantised model
    q_bn: str
    has_bias: bool = False
    w_bit: int = 4


def _unit_convs(cfg: NetConfig, stage: int, unit: int, inp: int, oup: int, stride: int) -> Dict[str, ConvSpec]:
    """Convs of one BaseNode; quantised names per QuantBaseNode.set_param (quant_modules.py:842-871)."""
    half = oup // 2
    r = "layer%d.%d." % (stage, unit)
    cin2 = inp if stride == 2 else half
    out = {}

    def add(tag, kind, cin, cout, k, s, groups, rc, rb, qn):
        out[tag] = ConvSpec(r + tag, kind, cin, cout, k, s, groups, r + rc, r + rb,
                            r + qn + ".conv", r + qn + ".bn", False, cfg.w_bit)

    add("pw1", "pw", cin2, half, 1, 1, 1, "b2.0", "b2.1", "quant_convbn1")
    add("dw2", "dw", half, half, 3, stride, half, "b2.3", "b2.4", "quant_convbn2")
    add("pw3", "pw", half, half, 1, 1, 1, "b2.5", "b2.6", "quant_convbn3")
    if stride == 2:
        add("dw4", "dw", inp, inp, 3, 2, inp, "b1.0", "b1.1", "quant_convbn4")
        add("pw5", "pw", inp, half, 1, 1, 1, "b1.2", "b1.3", "quant_convbn5")
    return out


@dataclass
class Graph:
    cfg: NetConfig
    stem: ConvSpec = None
    units: List[dict] = field(default_factory=list)      # {'stage','unit','stride','inp','oup','convs'}
    layer4: ConvSpec = None
    ups: List[dict] = field(default_factory=list)        # {'idx','cin','cout','scale','deform','channel'}
    heads: List[dict] = field(default_factory=list)      # {'name','classes','pw1','dw2','out'}

    def all_convs(self) -> List[ConvSpec]:
        out = [self.stem]
        for u in self.units:
            out.extend(u["convs"].values())
        out.append(self.layer4)
        for up in self.ups:
            out.extend([up["scale"], up["deform"], up["channel"]])
        for h in self.heads:
            out.extend([h["pw1"], h["dw2"], h["out"]])
        return out


def build_graph(cfg: NetConfig) -> Graph:
    g = Graph(cfg)
    ch = cfg.channels
    g.stem = ConvSpec("layer0", "stem", 3, ch[0], 3, 2 if cfg.maxpool else 4, 1,
                      "layer0.0", "layer0.1", "layer0.0.conv", "layer0.0.bn", False, 8)  # quantize_model.py:28
    for s, reps in enumerate(STAGE_REPEATS):
        inp, oup = ch[s], ch[s + 1]
        for u in range(reps + 1):
            stride = 2 if u == 0 else 1
            g.units.append(dict(stage=s + 1, unit=u, stride=stride, inp=inp, oup=oup,
                                convs=_unit_convs(cfg, s + 1, u, inp, oup, stride)))
    g.layer4 = ConvSpec("layer4", "pw", ch[3], ch[4], 1, 1, 1, "layer4.0", "layer4.1",
                        "layer4.0.conv", "layer4.0.bn", False, cfg.w_bit)
    for i, (cin, cout) in enumerate(zip(cfg.up_planes, UP_FILTERS)):
        raw = "deconv_layers.%d." % (4 * i)
        q = "deconv_layers.%d." % (3 * i)               # quantize_model.py:70-82 packs 3 modules per level
        g.ups.append(dict(
            idx=i, cin=cin, cout=cout,
            scale=ConvSpec("up%d.scale" % i, "scale", cin, 1, 1, 1, 1, raw + "conv_scale", "",
                           q + "quant_conv_scale", "", True, cfg.w_bit),
            deform=ConvSpec("up%d.deform" % i, "deform_dw", cin, cin, 3, 1, cin, raw + "conv", "",
                            q + "quant_deform_conv", "", False, cfg.w_bit),
            channel=ConvSpec("up%d.channel" % i, "pw", cin, cout, 1, 1, 1, raw + "conv_channel",
                             "deconv_layers.%d" % (4 * i + 1), q + "quant_conv_channel_bn.conv",
                             q + "quant_conv_channel_bn.bn", False, cfg.w_bit)))
    for name, classes in cfg.head_list():
        g.heads.append(dict(
            name=name, classes=classes,
            pw1=ConvSpec(name + ".pw1", "pw", 64, HEAD_CONV, 1, 1, 1, name + ".0", name + ".1",
                         name + ".quant_convbn1.conv", name + ".quant_convbn1.bn", False, cfg.w_bit),
            dw2=ConvSpec(name + ".dw2", "dw", HEAD_CONV, HEAD_CONV, 3, 1, HEAD_CONV, name + ".3", name + ".4",
                         name + ".quant_convbn2.conv", name + ".quant_convbn2.bn", False, cfg.w_bit),
            out=ConvSpec(name + ".out", "head_out", HEAD_CONV, classes, 1, 1, 1, name + ".6", "",
                         name + ".quant_conv", "", True, cfg.w_bit)))
    return g


# ---- activation-quantiser (QuantAct) key names in the quantised model -------------------------------------
def act_keys(g: Graph) -> Dict[str, str]:
    """Maps our label of every QuantAct to its state-dict prefix (buffers x_min / x_max, shape [1]).

    Stage-shared quantisers (quantize_model.py:40,51) appear once per unit in the state dict
    (`layerN.M.quant_act.*`), all aliasing one module; we read unit 0's copy.
    """
    k = {"stem": "layer0.1.1"}
    for u in g.units:
        r = "layer%d.%d." % (u["stage"], u["unit"])
        k[r + "act1"] = r + "quant_act1"
        k[r + "act2"] = r + "quant_act2"
        if u["stride"] == 2:
            k[r + "act4"] = r + "quant_act4"
        if u["unit"] == 0:
            k["layer%d.shared" % u["stage"]] = r + "quant_act"
    k["layer4"] = "layer4.1.1"
    for up in g.ups:
        q = "deconv_layers.%d." % (3 * up["idx"])
        k["up%d.s" % up["idx"]] = q + "quant_act.1"
        k["up%d.deform" % up["idx"]] = q + "quant_identity_deform"
        k["up%d.out" % up["idx"]] = "deconv_layers.%d.1" % (3 * up["idx"] + 1)
    for h in g.heads:
        k[h["name"] + ".act1"] = h["name"] + ".quant_act1.1"
        k[h["name"] + ".act3"] = h["name"] + ".quant_act3.1"
    return k


def raw_param_shapes(g: Graph) -> Dict[str, Tuple[int, ...]]:
    """Every tensor of the raw (pre-quantisation) state dict, in the reference's key space."""
    out = {}
    for c in g.all_convs():
        out[c.raw_conv + ".weight"] = (c.cout, c.cin // c.groups, c.k, c.k)
        if c.has_bias:
            out[c.raw_conv + ".bias"] = (c.cout,)
        if c.raw_bn:
            for f in ("weight", "bias", "running_mean", "running_var"):
                out[c.raw_bn + "." + f] = (c.cout,)
    return out


def raw_to_quant_key(g: Graph) -> Dict[str, str]:
    """raw key -> quantised key, for every parameter (what quantise-then-load does, base_detector.py:29-36)."""
    m = {}
    for c in g.all_convs():
        m[c.raw_conv + ".weight"] = c.q_conv + ".weight"
        if c.has_bias:
            m[c.raw_conv + ".bias"] = c.q_conv + ".bias"
        if c.raw_bn:
            for f in ("weight", "bias", "running_mean", "running_var"):
                m[c.raw_bn + "." + f] = c.q_bn + "." + f
    return m


def out_hw(cfg: NetConfig, h: int, w: int) -> Tuple[int, int]:
    """Spatial size of the head outputs (down ratio 4)."""
    return h // 4, w // 4
