"""Shared test-weight recipe.  TEST INFRASTRUCTURE ONLY (tests/, tests/golden/make_golden.py, __graft_entry__.smoke()).

At Keras-default initialisation the captioner is a degenerate test subject: the backbone's signal
decays below 1e-6 of its BatchNorm offsets within a few blocks, the co-attention softmax scales head outputs by 1/(H*W)
under LayerNorm's eps, and captions do not depend on the image.  `test_weights` keeps the reference's variable tree and
distributions (fpnmt.weights.init_weights) but
  * multiplies a few kernels by fixed gains so that every stage carries an O(1), input-dependent signal, and
  * sets every BatchNorm's moving statistics to the statistics of its input on four structured calibration images
    (`bn_calibration`), which is what training would have done.
Any float32 arrays are legitimate weights; parity only needs the oracle, the reference-under-shim run and the CUDA
engine to receive THE SAME arrays, which this function guarantees by being deterministic.
"""
from __future__ import annotations

import functools
from typing import Dict

import numpy as np
import torch

from .backbones import backbone_forward
from .ops import W, bn_calibration, nhwc_to_nchw

__all__ = ["TEST_GAINS", "CAPTION_GAINS", "test_weights", "caption_weights", "test_images", "bf16_conv_emulation"]

TEST_GAINS = {"/model/conv2d_4": 48.0, "/model/conv2d_5": 48.0, "pyramid_regression": 3.0, "pyramid_classification": 3.0,
              "final_layer": 6.0, "decoder/embedding": 20.0}


@functools.lru_cache(maxsize=8)
def _cached(backbone: str, vocab: int, layers: int, seed: int, cal_size: int):
    from fpnmt.synthetic import structured_images
    from fpnmt.weights import init_weights
    w = init_weights(backbone, vocab=vocab, seed=seed, num_layers=layers, randomize_bn=True, bias_std=0.02,
                     gains=TEST_GAINS)
    cal = torch.from_numpy(structured_images(4, cal_size, seed=99))
    with torch.no_grad(), bn_calibration():
        backbone_forward(backbone, nhwc_to_nchw(cal), W(w))
    return w


def test_weights(backbone: str, vocab: int = 512, layers: int = 2, seed: int = 0, cal_size: int = 256) -> Dict[str, np.ndarray]:
    """Deterministic test weights (a fresh dict; arrays shared, treat as read-only)."""
    return dict(_cached(backbone, vocab, layers, seed, cal_size))


test_weights.__test__ = False   # not a pytest test


# Decoder-side gains of `caption_weights` (on top of TEST_GAINS): a larger embedding and smaller cross-attention / FFN
# output projections raise the share of the decoder state that depends on the input token from ~2 % to ~50 %.
CAPTION_GAINS = {"decoder/embedding": 3.0, "mha2/dense": 0.3, "ffn2": 0.5}


@functools.lru_cache(maxsize=8)
def _caption_cached(backbone: str, vocab: int, layers: int, seed: int, cal_size: int, logit_std: float, end_bias: float):
    from .decode import transformer_logits
    from .model import create_look_ahead_mask, encoder
    w = dict(_cached(backbone, vocab, layers, seed, cal_size))
    for k in list(w):
        if k.startswith("transformer/decoder") and k.rsplit("/", 1)[1] in ("kernel", "embeddings"):
            for sub, g in CAPTION_GAINS.items():
                if sub in k:
                    w[k] = (w[k] * np.float32(g)).astype(np.float32)
    from fpnmt.synthetic import structured_images
    cal = torch.from_numpy(structured_images(4, cal_size, seed=98))
    t_cal = 16
    with torch.no_grad():
        mem = encoder(cal, W(w), backbone, num_layers=layers, input_vocab_size=(cal_size // 16) ** 2)
        tok = torch.randint(4, vocab, (4, t_cal), generator=torch.Generator().manual_seed(97))
        tok[:, 0] = 2
        lg, _ = transformer_logits(mem, tok, W(w), create_look_ahead_mask(t_cal), t_cal, num_layers=layers)
    lg = lg.reshape(-1, lg.shape[-1]).to(torch.float32)
    mean = lg.mean(0)
    s = np.float32(logit_std / float((lg - mean).std()))
    w["transformer/final_layer/kernel"] = (w["transformer/final_layer/kernel"] * s).astype(np.float32)
    b = ((w["transformer/final_layer/bias"] - mean.numpy()) * s).astype(np.float32)
    b[3] += np.float32(end_bias)          # <end> (id 3): how often a caption stops before max_len
    b[:3] -= np.float32(30.0)             # <pad>, <unk>, <start> never generated
    w["transformer/final_layer/bias"] = b
    return w


def caption_weights(backbone: str, vocab: int = 512, layers: int = 2, seed: int = 0, cal_size: int = 256,
                    logit_std: float = 5.0, end_bias: float = 0.0) -> Dict[str, np.ndarray]:
    """`test_weights` whose CAPTIONS are a meaningful test subject.

    With `test_weights` alone 98 % of the decoder's output state is one constant vector (cross-attention over a nearly
    uniform softmax + FFN bias paths), so every greedy caption is one token repeated and the reference's beam search
    and a true beam search coincide.  Here (i) CAPTION_GAINS raise the token-dependent share of the state and (ii) the
    final Dense layer is centred on a calibration set (bias -= mean logit, what training does within a few steps) and
    rescaled to a logit standard deviation of `logit_std`.  Measured on 100 structured images (MobileNetV2, 256x256,
    2 layers, V = 512, T = 16): 6..16 distinct tokens per caption, all captions different, `true_beam` leaves the greedy
    path on > 90 % of the images.  Deterministic; the same arrays go to the oracle and to the CUDA engine."""
    return dict(_caption_cached(backbone, vocab, layers, seed, cal_size, float(logit_std), float(end_bias)))


caption_weights.__test__ = False


def test_images(n: int, size: int = 256, seed: int = 1) -> torch.Tensor:
    from fpnmt.synthetic import structured_images
    return torch.from_numpy(structured_images(n, size, seed))


test_images.__test__ = False


class bf16_conv_emulation:
    """Context manager: every convolution of the oracle rounds its input, kernel and output to bfloat16 (fp32
    accumulation) — the arithmetic of the engine's BF16 fast mode.  Used to state that mode's stage tolerance relative
    to what bf16 rounding alone does to a random-init BatchNorm network (perturbations grow ~1.2x per layer)."""

    def __enter__(self):
        from . import backbones, model, ops
        self._mods = (ops, backbones, model)
        self._saved = [(m, n, getattr(m, n)) for m in self._mods for n in ("conv2d", "depthwise_conv2d") if hasattr(m, n)]
        oc, od = ops.conv2d, ops.depthwise_conv2d
        r = lambda t: t.to(torch.bfloat16).to(t.dtype)

        def conv2d(x, k, b=None, *a, **kw):
            return r(oc(r(x), r(k), b, *a, **kw))

        def depthwise_conv2d(x, k, *a, **kw):
            return r(od(r(x), k, *a, **kw))
        for m, n, _ in self._saved:
            setattr(m, n, conv2d if n == "conv2d" else depthwise_conv2d)
        return self

    def __exit__(self, *a):
        for m, n, f in self._saved:
            setattr(m, n, f)
        return False
