"""Primitive ops with TensorFlow/Keras semantics, restated on PyTorch CPU.  TEST INFRASTRUCTURE ONLY.

These stand in for the un-vendored tf.keras kernels the reference calls:
Conv2D / DepthwiseConv2D / ZeroPadding2D / MaxPooling2D / AveragePooling2D /
BatchNormalization / LayerNormalization / Dense / softmax.  CNN tensors are NCHW inside
this package (torch's conv layout); the public model functions take and return NHWC like
the reference.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence, Union

import numpy as np
import torch
import torch.nn.functional as F

__all__ = ["W", "bn_calibration", "tf_same_pad", "conv2d", "depthwise_conv2d", "max_pool", "avg_pool2", "batch_norm",
           "layer_norm", "dense", "leaky_relu", "relu6", "nhwc_to_nchw", "nchw_to_nhwc", "upsample_like", "load_image_array"]


class W:
    """Typed view over a weight dict (numpy float32) -> torch tensors of one dtype."""

    def __init__(self, weights: Dict[str, np.ndarray], dtype=torch.float32):
        self.w = weights
        self.dtype = dtype
        self._cache: Dict[str, torch.Tensor] = {}

    def __call__(self, key: str) -> torch.Tensor:
        t = self._cache.get(key)
        if t is None:
            t = torch.from_numpy(np.ascontiguousarray(self.w[key])).to(self.dtype)
            self._cache[key] = t
        return t

    def has(self, key: str) -> bool:
        return key in self.w


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 3, 1, 2).contiguous()


def nchw_to_nhwc(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 2, 3, 1).contiguous()


def tf_same_pad(size: int, k: int, stride: int, dilation: int = 1):
    """TensorFlow 'SAME': out = ceil(in/stride); extra pixel goes to the bottom/right."""
    out = -(-size // stride)
    keff = (k - 1) * dilation + 1
    total = max((out - 1) * stride + keff - size, 0)
    return total // 2, total - total // 2


def _pad(x: torch.Tensor, k, stride, padding, value=0.0):
    kh, kw = k
    if padding == "valid":
        return x
    if padding == "same":
        pt, pb = tf_same_pad(x.shape[2], kh, stride)
        pl, pr = tf_same_pad(x.shape[3], kw, stride)
    else:  # explicit ((top,bottom),(left,right)) == keras ZeroPadding2D
        (pt, pb), (pl, pr) = padding
    if pt or pb or pl or pr:
        x = F.pad(x, (pl, pr, pt, pb), value=value)
    return x


def conv2d(x: torch.Tensor, kernel_hwio: torch.Tensor, bias=None, stride: int = 1,
           padding: Union[str, Sequence] = "same") -> torch.Tensor:
    """tf.keras.layers.Conv2D on NCHW input with a Keras (kh,kw,Cin,Cout) kernel."""
    kh, kw = kernel_hwio.shape[0], kernel_hwio.shape[1]
    x = _pad(x, (kh, kw), stride, padding)
    wt = kernel_hwio.permute(3, 2, 0, 1).contiguous()
    return F.conv2d(x, wt, bias, stride=stride)


def depthwise_conv2d(x: torch.Tensor, kernel_hwc1: torch.Tensor, stride: int = 1,
                     padding: Union[str, Sequence] = "same") -> torch.Tensor:
    """tf.keras.layers.DepthwiseConv2D (depth_multiplier=1, no bias); kernel (kh,kw,C,1)."""
    kh, kw, c, _ = kernel_hwc1.shape
    x = _pad(x, (kh, kw), stride, padding)
    wt = kernel_hwc1.permute(2, 3, 0, 1).contiguous()   # (C,1,kh,kw)
    return F.conv2d(x, wt, None, stride=stride, groups=c)


def max_pool(x: torch.Tensor, k: int = 2, stride: int = 2, padding: Union[str, Sequence] = "valid") -> torch.Tensor:
    """tf.keras.layers.MaxPooling2D; 'same' pads with -inf (bottom/right first)."""
    if padding == "same":
        x = _pad(x, (k, k), stride, "same", value=float("-inf"))
    elif padding != "valid":
        x = _pad(x, (k, k), stride, padding, value=0.0)      # ZeroPadding2D + valid pool (DenseNet pool1)
    return F.max_pool2d(x, k, stride)


def avg_pool2(x: torch.Tensor) -> torch.Tensor:
    return F.avg_pool2d(x, 2, 2)


_CALIBRATE = [False]


class bn_calibration:
    """Context manager (test-weight preparation only): while active, every `batch_norm` call first overwrites the
    layer's moving_mean / moving_variance in the weight dict with the statistics of its actual input, so that a
    random-init backbone keeps an input-dependent signal of O(1) at every depth (at Keras-default init the signal of
    MobileNetV2 decays below 1e-6 of the BN offsets by C4, which would make stage-parity checks vacuous)."""

    def __enter__(self):
        _CALIBRATE[0] = True
        return self

    def __exit__(self, *a):
        _CALIBRATE[0] = False
        return False


def batch_norm(x: torch.Tensor, w: W, name: str, eps: float) -> torch.Tensor:
    """Inference-mode BatchNormalization over the channel axis of an NCHW tensor."""
    if _CALIBRATE[0]:
        xd = x.double()
        for leaf, val in (("moving_mean", xd.mean(dim=(0, 2, 3))), ("moving_variance", xd.var(dim=(0, 2, 3), unbiased=False))):
            w.w[name + "/" + leaf] = val.to(torch.float32).numpy().copy()
            w._cache.pop(name + "/" + leaf, None)
    g, b = w(name + "/gamma"), w(name + "/beta")
    m, v = w(name + "/moving_mean"), w(name + "/moving_variance")
    scale = g / torch.sqrt(v + eps)
    return x * scale.view(1, -1, 1, 1) + (b - m * scale).view(1, -1, 1, 1)


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """tf.keras.layers.LayerNormalization(axis=-1): biased variance, eps inside the sqrt."""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * gamma + beta


def dense(x: torch.Tensor, w: W, name: str) -> torch.Tensor:
    return x @ w(name + "/kernel") + w(name + "/bias")


def leaky_relu(x: torch.Tensor, alpha: float = 0.2) -> torch.Tensor:
    """tf.nn.leaky_relu default alpha=0.2 (common_definitions.py:14)."""
    return torch.where(x >= 0, x, x * alpha)


def relu6(x: torch.Tensor) -> torch.Tensor:
    return torch.clamp(x, 0.0, 6.0)


def upsample_like(src: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """layers/_misc.py:35-42 — nearest-neighbour resize of `src` to `target`'s HxW (NCHW here).

    TF2 `tf.image.resize(..., NEAREST_NEIGHBOR)` uses half-pixel centres: src index =
    floor((dst + 0.5) * in/out).  For the exact 2x upsamples on this path that is dst // 2
    (and equals the legacy TF1 convention too).
    """
    hin, win = src.shape[2], src.shape[3]
    hout, wout = target.shape[2], target.shape[3]
    iy = torch.clamp(torch.floor((torch.arange(hout, dtype=torch.float64) + 0.5) * (hin / hout)).long(), max=hin - 1)
    ix = torch.clamp(torch.floor((torch.arange(wout, dtype=torch.float64) + 0.5) * (win / wout)).long(), max=win - 1)
    return src[:, :, iy][:, :, :, ix]


def load_image_array(img_hwc_u8: np.ndarray, size: int = 512) -> np.ndarray:
    """dataset.py:19-26 after `tf.image.decode_jpeg`: `tf.image.resize(img, (size, size))` (TF2 default: bilinear with
    half-pixel centres, antialias=False; source index = (dst + 0.5) * scale - 0.5, neighbours clamped to the image) and
    `mobilenet_v2.preprocess_input` (x / 127.5 - 1).  uint8 (H,W,3) -> float32 (size,size,3) in [-1, 1]."""
    a = np.asarray(img_hwc_u8).astype(np.float32)
    h, w = a.shape[0], a.shape[1]

    def axis(n_in, n_out):
        s = (np.arange(n_out, dtype=np.float32) + np.float32(0.5)) * np.float32(n_in / n_out) - np.float32(0.5)
        lo = np.floor(s)
        return (np.clip(lo, 0, n_in - 1).astype(np.int64), np.clip(lo + 1, 0, n_in - 1).astype(np.int64),
                (s - lo).astype(np.float32))
    y0, y1, fy = axis(h, size)
    x0, x1, fx = axis(w, size)
    fx_, fy_ = fx[None, :, None], fy[:, None, None]
    top = a[y0][:, x0] * (1 - fx_) + a[y0][:, x1] * fx_
    bot = a[y1][:, x0] * (1 - fx_) + a[y1][:, x1] * fx_
    return ((top * (1 - fy_) + bot * fy_) / np.float32(127.5) - np.float32(1.0)).astype(np.float32)
