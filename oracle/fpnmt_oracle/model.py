"""FPN + heads + Multi-Transformer, restated from the reference.  TEST INFRASTRUCTURE ONLY.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from .backbones import RN, backbone_forward
from .ops import (W, conv2d, dense, layer_norm, leaky_relu, max_pool, nchw_to_nhwc, nhwc_to_nchw, upsample_like)

__all__ = ["get_angles", "raw_positional_encoding", "positional_encoding", "create_look_ahead_mask",
           "scaled_dot_product_attention", "mha", "coattention_cnn", "pyramid_features", "head_trunk",
           "feature_extractor", "encoder_tokens", "encoder_layer", "encoder", "decoder_layer", "decoder",
           "transformer_logits", "FE", "HM", "TR", "X_ORDER", "NUM_OF_PYRAMIDS", "BASELINE_INDEX"]

TR = "transformer"
FE = TR + "/encoder/feature_extractor"
HM = FE + "/model"

NUM_OF_PYRAMIDS = 5          # common/common_definitions.py:66
BASELINE_INDEX = 3           # common/common_definitions.py:70
N_CONV_SUBMODULE = 2         # common/common_definitions.py:67
X_ORDER = [i for i in range(NUM_OF_PYRAMIDS) if i != BASELINE_INDEX] + [BASELINE_INDEX]   # transformer.py:253


# ---- positional encoding & masks (models/transformer.py:22-43, 54-56) ----------------------------
def get_angles(pos, i, d_model):
    angle_rates = 1 / np.power(10000, (2 * (i // 2)) / np.float32(d_model))      # transformer.py:23
    return pos * angle_rates


def raw_positional_encoding(position: int, d_model: int) -> torch.Tensor:
    angle_rads = get_angles(np.arange(position)[:, np.newaxis], np.arange(d_model)[np.newaxis, :], d_model)
    angle_rads[:, 0::2] = np.sin(angle_rads[:, 0::2])                             # transformer.py:34
    angle_rads[:, 1::2] = np.cos(angle_rads[:, 1::2])                             # transformer.py:37
    return torch.from_numpy(angle_rads.astype(np.float32))                        # tf.cast(..., float32)


def positional_encoding(position: int, d_model: int) -> torch.Tensor:
    return raw_positional_encoding(position, d_model)[None]


def create_look_ahead_mask(size: int, dtype=torch.float32) -> torch.Tensor:
    return 1 - torch.tril(torch.ones(size, size, dtype=dtype))                    # transformer.py:55


# ---- attention (models/transformer.py:70-155) ------------------------------------------------------
def scaled_dot_product_attention(q, k, v, mask):
    matmul_qk = q @ k.transpose(-1, -2)                                           # transformer.py:88
    dk = torch.tensor(float(k.shape[-1]), dtype=q.dtype)
    logits = matmul_qk / torch.sqrt(dk)                                           # transformer.py:91-92
    if mask is not None:
        logits = logits + mask.to(q.dtype) * -1e9                                 # transformer.py:95-96
    weights = torch.softmax(logits, dim=-1)                                       # transformer.py:100
    return weights @ v, weights


def mha(w: W, name: str, v, k, q, mask, num_heads: int = 8):
    """MultiHeadAttention.call (transformer.py:131-155)."""
    bsz, d_model = q.shape[0], q.shape[-1]
    depth = d_model // num_heads
    q = dense(q, w, name + "/wq")
    k = dense(k, w, name + "/wk")
    v = dense(v, w, name + "/wv")

    def split(x):                                                                 # transformer.py:124-129
        return x.reshape(bsz, -1, num_heads, depth).permute(0, 2, 1, 3)

    att, weights = scaled_dot_product_attention(split(q), split(k), split(v), mask)
    att = att.permute(0, 2, 1, 3).reshape(bsz, -1, d_model)                       # transformer.py:147-151
    return dense(att, w, name + "/dense"), weights


# ---- CNN side (models/retinanet.py, models/coattention.py, layers/_misc.py) -----------------------
def coattention_cnn(score: torch.Tensor, hs: torch.Tensor) -> torch.Tensor:
    """CoAttention_CNN.call (coattention.py:13-32) on NCHW tensors: softmax over all H*W positions."""
    b = score.shape[0]
    wts = torch.softmax(score.reshape(b, -1), dim=1).reshape(score.shape)          # coattention.py:24-27
    return wts * hs                                                                # coattention.py:30


def _conv_named(x, w: W, name: str, act: Optional[str] = None):
    y = conv2d(x, w(name + "/kernel"), w(name + "/bias"), 1, "same")
    if act == "relu":
        y = torch.relu(y)
    elif act == "leaky":
        y = leaky_relu(y, 0.2)
    return y


def pyramid_features(c3, c4, c5, w: W) -> List[torch.Tensor]:
    """__create_pyramid_features (retinanet.py:105-141).  Returns [P3,P4,P5,P6,P7] (NCHW)."""
    p5_feat = _conv_named(c5, w, RN + "/C5_reduced")                               # :118
    p5_up = upsample_like(p5_feat, c4)                                             # :119
    p5 = _conv_named(p5_feat, w, RN + "/P5", "relu")                               # :120
    p4 = _conv_named(c4, w, RN + "/C4_reduced")                                    # :123
    p4 = p5_up + p4                                                                # :124 P4_merged
    p4_up = upsample_like(p4, c3)                                                  # :125 (from the merged, pre-3x3 map)
    p4 = _conv_named(p4, w, RN + "/P4", "relu")                                    # :126
    p3 = _conv_named(c3, w, RN + "/C3_reduced")                                    # :129
    p3 = p4_up + p3                                                                # :130
    p3 = _conv_named(p3, w, RN + "/P3", "relu")                                    # :131
    p6 = max_pool(_conv_named(p5_feat, w, RN + "/conv2d", "relu"), 2, 2, "valid")  # :134-135
    p7 = max_pool(_conv_named(p6, w, RN + "/conv2d_1", "relu"), 2, 2, "valid")     # :138-139
    return [p3, p4, p5, p6, p7]


def head_trunk(x, w: W, sub: str, layer: str):
    """default_{regression,classification}_model truncated at layers[N_CONV_SUBMODULE] (retinanet.py:283-284)."""
    for i in range(N_CONV_SUBMODULE):
        x = _conv_named(x, w, RN + "/%s/%s_%d" % (sub, layer, i), "relu")
    return x


def head(p, w: W):
    """Per-level sub-model built in FeatureExtractor.__init__ (retinanet.py:283-297)."""
    reg = head_trunk(p, w, "regression_submodel", "pyramid_regression")
    cls = head_trunk(p, w, "classification_submodel", "pyramid_classification")
    regression = _conv_named(reg, w, HM + "/conv2d_2")                             # :287 linear, 256->1
    classification = _conv_named(cls, w, HM + "/conv2d_3")                         # :288 linear, 256->256
    out = coattention_cnn(regression, classification)                              # :291
    out = _conv_named(out, w, HM + "/conv2d_4", "leaky")                           # :292
    out = max_pool(out, 2, 2, "valid")                                             # :293
    return _conv_named(out, w, HM + "/conv2d_5", "leaky")                          # :294


def feature_extractor(images_nhwc: torch.Tensor, w: W, backbone: str = "mobilenet224_1.0",
                      taps: Optional[Dict[str, torch.Tensor]] = None) -> List[torch.Tensor]:
    """FeatureExtractor.call (retinanet.py:306-307): list of 5 NHWC maps for P3..P7."""
    x = nhwc_to_nchw(images_nhwc.to(w.dtype))
    c3, c4, c5 = backbone_forward(backbone, x, w)
    ps = pyramid_features(c3, c4, c5, w)
    outs = [head(p, w) for p in ps]                                                # :300-301
    if taps is not None:
        for n, t in zip(("C3", "C4", "C5"), (c3, c4, c5)):
            taps[n] = nchw_to_nhwc(t)
        for n, t in zip(("P3", "P4", "P5", "P6", "P7"), ps):
            taps[n] = nchw_to_nhwc(t)
    return [nchw_to_nhwc(o) for o in outs]


# ---- Multi-Transformer encoder (models/transformer.py:158-200, 246-303) ---------------------------
def encoder_tokens(features_nhwc: List[torch.Tensor], w: W, pos_encoding: torch.Tensor) -> List[torch.Tensor]:
    """Encoder.call pre-amble (transformer.py:279-296): reorder, flatten, shared LN, + pos-enc."""
    x = [features_nhwc[i] for i in X_ORDER]                                        # :279
    out = []
    g, b = w(TR + "/encoder/layernorm1/gamma"), w(TR + "/encoder/layernorm1/beta")
    for _x in x:
        bsz, h, wd, c = _x.shape
        seq_len = h * wd
        _x = _x.reshape(bsz, seq_len, c)                                           # :289
        _x = layer_norm(_x, g, b, 1e-6)                                            # :290
        _x = _x + pos_encoding[:, :seq_len, :].to(_x.dtype)                        # :292
        out.append(_x)
    return out


def encoder_layer(x: List[torch.Tensor], w: W, name: str, num_heads: int = 8) -> torch.Tensor:
    """EncoderLayer.call (transformer.py:176-200), training=False, mask=None."""
    baseline = x[NUM_OF_PYRAMIDS - 1]
    out = baseline
    for i in range(NUM_OF_PYRAMIDS - 1):
        m, _ = mha(w, name + "/mhas/%d" % i, x[i], x[i], baseline, None, num_heads)   # :187
        out = out + m                                                                # :188
    out1 = layer_norm(out, w(name + "/layernorm1/gamma"), w(name + "/layernorm1/beta"), 1e-6)
    ffn = leaky_relu(dense(out1, w, name + "/ffn1"), 0.2)                           # :194
    ffn = dense(ffn, w, name + "/ffn2")                                             # :195
    return layer_norm(out1 + ffn, w(name + "/layernorm2/gamma"), w(name + "/layernorm2/beta"), 1e-6)


def encoder(images_nhwc: torch.Tensor, w: W, backbone: str = "mobilenet224_1.0", num_layers: int = 6,
            num_heads: int = 8, input_vocab_size: int = 1024, taps: Optional[dict] = None) -> torch.Tensor:
    """Encoder.call (transformer.py:266-303) -> (B, 16, d_model)."""
    feats = feature_extractor(images_nhwc, w, backbone, taps)
    d_model = feats[0].shape[-1]
    pos = positional_encoding(input_vocab_size, d_model)                            # transformer.py:255
    x = encoder_tokens(feats, w, pos)
    if taps is not None:
        taps["features"] = feats
        taps["tokens"] = [t.clone() for t in x]
    for l in range(num_layers):
        x[NUM_OF_PYRAMIDS - 1] = encoder_layer(x, w, TR + "/encoder/enc_layers/%d" % l, num_heads)   # :298-299
        if taps is not None:
            taps["enc_layer%d" % l] = x[NUM_OF_PYRAMIDS - 1].clone()
    return x[NUM_OF_PYRAMIDS - 1]


# ---- decoder (models/transformer.py:203-243, 306-341, 359-374) ------------------------------------
def decoder_layer(x, enc_output, w: W, name: str, look_ahead_mask, padding_mask=None, num_heads: int = 8):
    """DecoderLayer.call (transformer.py:224-243), training=False."""
    attn1, w1 = mha(w, name + "/mha1", x, x, x, look_ahead_mask, num_heads)          # :228
    out1 = layer_norm(attn1 + x, w(name + "/layernorm1/gamma"), w(name + "/layernorm1/beta"), 1e-6)
    attn2, w2 = mha(w, name + "/mha2", enc_output, enc_output, out1, padding_mask, num_heads)   # :232-233
    out2 = layer_norm(attn2 + out1, w(name + "/layernorm2/gamma"), w(name + "/layernorm2/beta"), 1e-6)
    ffn = dense(leaky_relu(dense(out2, w, name + "/ffn1"), 0.2), w, name + "/ffn2")  # :237-238
    out3 = layer_norm(ffn + out2, w(name + "/layernorm3/gamma"), w(name + "/layernorm3/beta"), 1e-6)
    return out3, w1, w2


def decoder(tokens: torch.Tensor, enc_output, w: W, look_ahead_mask, max_seq_len: int, num_layers: int = 6,
            num_heads: int = 8, max_position: int = 0, taps: Optional[dict] = None):
    """Decoder.call (transformer.py:321-341).  tokens: (B, t) int64."""
    seq_len = tokens.shape[1]
    emb = w(TR + "/decoder/embedding/embeddings")
    d_model = emb.shape[1]
    x = emb[tokens]                                                                 # :326 (no sqrt(d) scaling, :327)
    pos = raw_positional_encoding(max_seq_len + max_position, d_model).to(x.dtype)  # :315
    x = x + pos[None, :seq_len, :]                                                  # :329
    attn = {}
    for i in range(num_layers):
        x, b1, b2 = decoder_layer(x, enc_output, w, TR + "/decoder/dec_layers/%d" % i, look_ahead_mask, None, num_heads)
        attn["decoder_layer%d_block1" % (i + 1)] = b1                               # :337-338
        attn["decoder_layer%d_block2" % (i + 1)] = b2
        if taps is not None:
            taps["dec_layer%d" % i] = x.clone()
    return x, attn


def transformer_logits(enc_output, tokens, w: W, look_ahead_mask, max_seq_len: int, num_layers: int = 6,
                       num_heads: int = 8, taps: Optional[dict] = None):
    """Transformer.call, training=False branch (transformer.py:359-374): inp is the encoder output."""
    dec_output, attn = decoder(tokens, enc_output, w, look_ahead_mask, max_seq_len, num_layers, num_heads, 0, taps)
    return dense(dec_output, w, TR + "/final_layer"), attn                          # :372
