"""Variable tree of the captioner and a deterministic random-init factory.

The tree follows the reference's `tf.train.Checkpoint(transformer=...)` object graph
(/root/reference/utils/pipeline.py:38-39): attribute names and list indices of
`Transformer` (/root/reference/models/transformer.py:344-357), `Encoder` (:246-264),
`EncoderLayer` (:158-174), `Decoder` (:306-319), `DecoderLayer` (:203-222),
`MultiHeadAttention` (:107-122) and `FeatureExtractor`
(/root/reference/models/retinanet.py:266-304).  Layouts are Keras': Dense kernel
(in,out), Conv2D kernel (kh,kw,Cin,Cout), DepthwiseConv2D kernel (kh,kw,C,1),
BatchNormalization {gamma,beta,moving_mean,moving_variance}, LayerNormalization
{gamma,beta}, Embedding {embeddings}.

Backbone layer names are the upstream ones the reference depends on
(mobilenet.py:64 `block_5_add`/`block_12_add`/`out_relu`; resnet.py:99,112;
densenet.py:93-96) — see SURVEY.md Appendix C for provenance.

A weight set is a plain `dict[str, np.ndarray(float32)]`; it can be stored as .npz.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Tuple

import numpy as np

from . import config as C

TR = "transformer"
FE = TR + "/encoder/feature_extractor"
RN = FE + "/retinanet_model"
HM = FE + "/model"

Spec = Tuple[str, str, tuple, str]   # (key prefix, kind, shape, initializer)


# --------------------------------------------------------------------------------------
# architecture specs (names + shapes only; no arithmetic)
# --------------------------------------------------------------------------------------
def _conv(name, kh, kw, cin, cout, init, bias):
    out = [(name + "/kernel", (kh, kw, cin, cout), init)]
    if bias:
        out.append((name + "/bias", (cout,), "zeros"))
    return out


def _bn(name, c):
    return [(name + "/gamma", (c,), "ones"), (name + "/beta", (c,), "zeros"),
            (name + "/moving_mean", (c,), "zeros"), (name + "/moving_variance", (c,), "ones")]


def _dense(name, cin, cout, init):
    return [(name + "/kernel", (cin, cout), init), (name + "/bias", (cout,), "zeros")]


def _ln(name, c):
    return [(name + "/gamma", (c,), "ones"), (name + "/beta", (c,), "zeros")]


MOBILENETV2_BLOCKS = [  # (block id, cout, stride) ; Keras MobileNetV2 alpha=1.0
    (1, 24, 2), (2, 24, 1), (3, 32, 2), (4, 32, 1), (5, 32, 1), (6, 64, 2), (7, 64, 1),
    (8, 64, 1), (9, 64, 1), (10, 96, 1), (11, 96, 1), (12, 96, 1), (13, 160, 2),
    (14, 160, 1), (15, 160, 1), (16, 320, 1)]

RESNET50_STAGES = [3, 4, 6, 3]
DENSENET121_BLOCKS = [6, 12, 24, 16]


def mobilenetv2_spec() -> List[tuple]:
    s = []
    g = "glorot_uniform"
    s += _conv(RN + "/Conv1", 3, 3, 3, 32, g, False) + _bn(RN + "/bn_Conv1", 32)
    s += [(RN + "/expanded_conv_depthwise/depthwise_kernel", (3, 3, 32, 1), g)]
    s += _bn(RN + "/expanded_conv_depthwise_BN", 32)
    s += _conv(RN + "/expanded_conv_project", 1, 1, 32, 16, g, False) + _bn(RN + "/expanded_conv_project_BN", 16)
    cin = 16
    for k, cout, _stride in MOBILENETV2_BLOCKS:
        p = RN + "/block_%d" % k
        s += _conv(p + "_expand", 1, 1, cin, 6 * cin, g, False) + _bn(p + "_expand_BN", 6 * cin)
        s += [(p + "_depthwise/depthwise_kernel", (3, 3, 6 * cin, 1), g)] + _bn(p + "_depthwise_BN", 6 * cin)
        s += _conv(p + "_project", 1, 1, 6 * cin, cout, g, False) + _bn(p + "_project_BN", cout)
        cin = cout
    s += _conv(RN + "/Conv_1", 1, 1, 320, 1280, g, False) + _bn(RN + "/Conv_1_bn", 1280)
    return s


def resnet50_spec() -> List[tuple]:
    s = []
    h = "he_normal"
    s += _conv(RN + "/conv1", 7, 7, 3, 64, h, False) + _bn(RN + "/bn_conv1", 64)
    cin = 64
    for st, nblk in enumerate(RESNET50_STAGES):
        f = 64 * 2 ** st
        for b in range(nblk):
            nm = "%d%s" % (st + 2, chr(ord("a") + b))
            s += _conv(RN + "/res%s_branch2a" % nm, 1, 1, cin, f, h, False) + _bn(RN + "/bn%s_branch2a" % nm, f)
            s += _conv(RN + "/res%s_branch2b" % nm, 3, 3, f, f, h, False) + _bn(RN + "/bn%s_branch2b" % nm, f)
            s += _conv(RN + "/res%s_branch2c" % nm, 1, 1, f, 4 * f, h, False) + _bn(RN + "/bn%s_branch2c" % nm, 4 * f)
            if b == 0:
                s += _conv(RN + "/res%s_branch1" % nm, 1, 1, cin, 4 * f, h, False) + _bn(RN + "/bn%s_branch1" % nm, 4 * f)
            cin = 4 * f
    return s


def densenet121_spec() -> List[tuple]:
    s = []
    g = "glorot_uniform"
    s += _conv(RN + "/conv1/conv", 7, 7, 3, 64, g, False) + _bn(RN + "/conv1/bn", 64)
    c = 64
    for si, nblk in enumerate(DENSENET121_BLOCKS):
        stage = si + 2
        for b in range(1, nblk + 1):
            p = RN + "/conv%d_block%d" % (stage, b)
            s += _bn(p + "_0_bn", c) + _conv(p + "_1_conv", 1, 1, c, 128, g, False)
            s += _bn(p + "_1_bn", 128) + _conv(p + "_2_conv", 3, 3, 128, 32, g, False)
            c += 32
        if si < 3:
            p = RN + "/pool%d" % stage
            s += _bn(p + "_bn", c) + _conv(p + "_conv", 1, 1, c, c // 2, g, False)
            c //= 2
    return s


BACKBONE_TAPS = {  # channels of C3, C4, C5
    "mobilenet224_1.0": (32, 96, 1280),
    "resnet50": (512, 1024, 2048),
    "densenet121": (512, 1024, 1024),
}


def backbone_spec(backbone: str) -> List[tuple]:
    if backbone.startswith("mobilenet"):
        return mobilenetv2_spec()
    if backbone == "resnet50":
        return resnet50_spec()
    if backbone == "densenet121":
        return densenet121_spec()
    raise ValueError("Backbone ('%s') is invalid." % backbone)


def fpn_head_spec(backbone: str) -> List[tuple]:
    """retinanet.py:105-141 (FPN), :25-102 (trunks), :287-294 (head convs)."""
    c3, c4, c5 = BACKBONE_TAPS[backbone]
    f = C.NUM_OF_RETINANET_FILTERS
    g, n01, h = "glorot_uniform", "normal0.01", "he_normal"
    s = []
    s += _conv(RN + "/C5_reduced", 1, 1, c5, f, g, True) + _conv(RN + "/P5", 3, 3, f, f, g, True)
    s += _conv(RN + "/C4_reduced", 1, 1, c4, f, g, True) + _conv(RN + "/P4", 3, 3, f, f, g, True)
    s += _conv(RN + "/C3_reduced", 1, 1, c3, f, g, True) + _conv(RN + "/P3", 3, 3, f, f, g, True)
    s += _conv(RN + "/conv2d", 3, 3, f, f, g, True) + _conv(RN + "/conv2d_1", 3, 3, f, f, g, True)
    for i in range(C.N_CONV_SUBMODULE):
        s += _conv(RN + "/regression_submodel/pyramid_regression_%d" % i, 3, 3, f, f, n01, True)
    for i in range(C.N_CONV_SUBMODULE):
        s += _conv(RN + "/classification_submodel/pyramid_classification_%d" % i, 3, 3, f, f, n01, True)
    s += _conv(HM + "/conv2d_2", 3, 3, f, 1, h, True)          # regression score   (retinanet.py:287)
    s += _conv(HM + "/conv2d_3", 3, 3, f, f, h, True)          # classification map (retinanet.py:288)
    s += _conv(HM + "/conv2d_4", 3, 3, f, f, h, True)          # after co-attention (retinanet.py:292)
    s += _conv(HM + "/conv2d_5", 3, 3, f, C.d_model, h, True)  # after max-pool     (retinanet.py:294)
    return s


def _mha(name, d):
    h = "he_normal"
    return _dense(name + "/wq", d, d, h) + _dense(name + "/wk", d, d, h) + _dense(name + "/wv", d, d, h) \
        + _dense(name + "/dense", d, d, h)


def transformer_spec(num_layers: int, d: int, dff: int, vocab: int) -> List[tuple]:
    h = "he_normal"
    s = _ln(TR + "/encoder/layernorm1", d)
    for l in range(num_layers):
        p = TR + "/encoder/enc_layers/%d" % l
        for i in range(C.NUM_OF_PYRAMIDS - 1):
            s += _mha(p + "/mhas/%d" % i, d)
        s += _dense(p + "/ffn1", d, dff, h) + _dense(p + "/ffn2", dff, d, h)
        s += _ln(p + "/layernorm1", d) + _ln(p + "/layernorm2", d)
    s += [(TR + "/decoder/embedding/embeddings", (vocab, d), "uniform0.05")]
    for l in range(num_layers):
        p = TR + "/decoder/dec_layers/%d" % l
        s += _mha(p + "/mha1", d) + _mha(p + "/mha2", d)
        s += _dense(p + "/ffn1", d, dff, h) + _dense(p + "/ffn2", dff, d, h)
        s += _ln(p + "/layernorm1", d) + _ln(p + "/layernorm2", d) + _ln(p + "/layernorm3", d)
    s += [(TR + "/final_layer/kernel", (d, vocab), "glorot_uniform"), (TR + "/final_layer/bias", (vocab,), "zeros")]
    return s


def model_spec(backbone: str, vocab: int, num_layers: int = C.num_layers, d: int = C.d_model,
               dff: int = C.dff) -> List[tuple]:
    return backbone_spec(backbone) + fpn_head_spec(backbone) + transformer_spec(num_layers, d, dff, vocab)


# --------------------------------------------------------------------------------------
# initialisers (Keras semantics)
# --------------------------------------------------------------------------------------
def _fans(shape):
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = int(np.prod(shape[:-2]))
    return shape[-2] * rf, shape[-1] * rf


def _draw(rng: np.random.Generator, shape, init: str) -> np.ndarray:
    if init == "zeros":
        return np.zeros(shape, np.float32)
    if init == "ones":
        return np.ones(shape, np.float32)
    if init == "normal0.01":
        return rng.normal(0.0, 0.01, shape).astype(np.float32)
    if init == "uniform0.05":
        return rng.uniform(-0.05, 0.05, shape).astype(np.float32)
    fan_in, fan_out = _fans(shape)
    if init == "glorot_uniform":
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-lim, lim, shape).astype(np.float32)
    if init == "he_normal":
        std = math.sqrt(2.0 / fan_in) / 0.87962566103423978
        x = rng.normal(0.0, 1.0, shape)
        bad = np.abs(x) > 2.0
        while bad.any():                       # truncated normal at +-2 sigma by resampling
            x[bad] = rng.normal(0.0, 1.0, int(bad.sum()))
            bad = np.abs(x) > 2.0
        return (x * std).astype(np.float32)
    raise ValueError(init)


def init_weights(backbone: str = "mobilenet224_1.0", vocab: int = C.SYNTH_VOCAB, seed: int = 0,
                 num_layers: int = C.num_layers, d: int = C.d_model, dff: int = C.dff,
                 randomize_bn: bool = False, bias_std: float = 0.0,
                 gains: Dict[str, float] | None = None) -> Dict[str, np.ndarray]:
    """Random weights with the reference's initial distributions.

    randomize_bn : draw BN moving stats / affine from non-trivial ranges so that BN folding is tested.
    bias_std     : draw every bias from N(0, bias_std) instead of zeros so that bias paths are tested.
    gains        : {key substring: factor} multiplied into matching *kernel* arrays (test-only knob to
                   keep signal magnitudes O(1); see SURVEY.md §7.2 "random-init signal scale").
    """
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}
    for key, shape, init in model_spec(backbone, vocab, num_layers, d, dff):
        a = _draw(rng, shape, init)
        leaf = key.rsplit("/", 1)[1]
        if randomize_bn:
            if leaf == "moving_mean":
                a = rng.normal(0.0, 0.1, shape).astype(np.float32)
            elif leaf == "moving_variance":
                a = rng.uniform(0.5, 1.5, shape).astype(np.float32)
            elif leaf == "gamma" and "layernorm" not in key:
                a = rng.uniform(0.8, 1.2, shape).astype(np.float32)
            elif leaf == "beta" and "layernorm" not in key:
                a = rng.normal(0.0, 0.1, shape).astype(np.float32)
        if bias_std > 0.0 and leaf == "bias":
            a = rng.normal(0.0, bias_std, shape).astype(np.float32)
        if bias_std > 0.0 and "layernorm" in key:
            a = (a + rng.normal(0.0, bias_std, shape)).astype(np.float32)
        if gains and leaf in ("kernel", "depthwise_kernel", "embeddings"):
            for sub, gfac in gains.items():
                if sub in key:
                    a = (a * np.float32(gfac)).astype(np.float32)
        w[key] = a
    return w


def save_weights(path: str, w: Dict[str, np.ndarray]) -> None:
    np.savez(path, **{k.replace("/", "|"): v for k, v in w.items()})


def load_weights(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        return {k.replace("|", "/"): z[k] for k in z.files}


def param_count(w: Dict[str, np.ndarray]) -> int:
    return int(sum(v.size for v in w.values()))
