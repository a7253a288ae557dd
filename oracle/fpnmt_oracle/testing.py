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

__all__ = ["TEST_GAINS", "test_weights", "test_images", "bf16_conv_emulation"]

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
