"""Host-side mirror of /root/reference/models/retinanet.py (+ the backbone builder modules).

`FeatureExtractor(retinanet_weight_path=None)` / `__call__(inp) -> list of 5 maps` (retinanet.py:266-307) and the
builder signatures `retinanet(inputs, backbone_layers, num_classes, ...)` (retinanet.py:217-225),
`mobilenet_retinanet` (mobilenet.py:43), `resnet_retinanet` / `resnet50_retinanet` (resnet.py:78,115),
`densenet_retinanet` (densenet.py:73) and the `backbone(name)` factory (models/__init__.py:49-63) are kept.
The reference builders return Keras functional models; here they return a `RetinaNetSpec` describing the same
graph (backbone name + C3/C4/C5 tap names), which `FeatureExtractor` hands to the CUDA engine.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import config as C
from .weights import BACKBONE_TAPS, init_weights, load_weights

BACKBONE_LAYER_NAMES = {
    "mobilenet224_1.0": ("block_5_add", "block_12_add", "out_relu"),                       # mobilenet.py:64
    "resnet50": ("res3d_relu", "res4f_relu", "res5c_relu"),                                # resnet.py:112 outputs[1:]
    "densenet121": ("conv3_block12_concat", "conv4_block24_concat", "conv5_block16_concat"),  # densenet.py:93-96
}


@dataclass
class RetinaNetSpec:
    backbone: str
    backbone_layers: Tuple[str, str, str]
    num_classes: int = C.NUM_OF_CLASSES
    num_anchors: int = C.NUM_OF_ANCHORS
    name: str = "retinanet"

    @property
    def tap_channels(self):
        return BACKBONE_TAPS[self.backbone]


def retinanet(inputs, backbone_layers, num_classes, num_anchors=None, create_pyramid_features=None, submodels=None,
              name="retinanet", *, backbone: str = "mobilenet224_1.0") -> RetinaNetSpec:
    """retinanet.py:217-263.  Custom `create_pyramid_features` / `submodels` functors are not supported: the
    engine implements the reference's own FPN (retinanet.py:105-141) and default sub-models (:25-102)."""
    if create_pyramid_features is not None or submodels is not None:
        raise NotImplementedError("only the reference's default pyramid and sub-models are implemented in CUDA")
    if num_anchors is None:
        num_anchors = C.NUM_OF_ANCHORS
    return RetinaNetSpec(backbone, tuple(backbone_layers), num_classes, num_anchors, name)


def mobilenet_retinanet(num_classes, backbone="mobilenet224_1.0", inputs=None, modifier=None, **kwargs) -> RetinaNetSpec:
    """mobilenet.py:43-72."""
    alpha = float(backbone.split("_")[1])
    if alpha != 1.0 or not backbone.startswith("mobilenet224"):
        raise ValueError("Backbone ('{}') not implemented: only mobilenet224_1.0 is wired (retinanet.py:274)".format(backbone))
    if modifier:
        raise NotImplementedError("backbone modifiers are a training-time feature")
    return retinanet(inputs=inputs, num_classes=num_classes, backbone_layers=BACKBONE_LAYER_NAMES[backbone],
                     backbone=backbone, **kwargs)


def resnet_retinanet(num_classes, backbone="resnet50", inputs=None, modifier=None, **kwargs) -> RetinaNetSpec:
    """resnet.py:78-112."""
    if backbone != "resnet50":
        raise ValueError("Backbone ('{}') is invalid.".format(backbone))
    if modifier:
        raise NotImplementedError("backbone modifiers are a training-time feature")
    return retinanet(inputs=inputs, num_classes=num_classes, backbone_layers=BACKBONE_LAYER_NAMES[backbone],
                     backbone=backbone, **kwargs)


def resnet50_retinanet(num_classes, inputs=None, **kwargs) -> RetinaNetSpec:
    """resnet.py:115-116."""
    return resnet_retinanet(num_classes=num_classes, backbone="resnet50", inputs=inputs, **kwargs)


def densenet_retinanet(num_classes, backbone="densenet121", inputs=None, modifier=None, **kwargs) -> RetinaNetSpec:
    """densenet.py:73-105."""
    if backbone != "densenet121":
        raise ValueError("Backbone ('{}') not in allowed backbones (densenet121).".format(backbone))
    if modifier:
        raise NotImplementedError("backbone modifiers are a training-time feature")
    return retinanet(inputs=inputs, num_classes=num_classes, backbone_layers=BACKBONE_LAYER_NAMES[backbone],
                     backbone=backbone, **kwargs)


class Backbone:
    """models/__init__.py:5-46 (the detection-only custom_objects are out of scope)."""

    def __init__(self, backbone):
        self.backbone = backbone
        self.validate()

    def validate(self):
        if self.backbone not in BACKBONE_LAYER_NAMES:
            raise ValueError("Backbone ('{}') not in allowed backbones ({}).".format(self.backbone, list(BACKBONE_LAYER_NAMES)))

    def retinanet(self, *args, **kwargs):
        if self.backbone.startswith("mobilenet"):
            return mobilenet_retinanet(*args, backbone=self.backbone, **kwargs)
        if "resnet" in self.backbone:
            return resnet_retinanet(*args, backbone=self.backbone, **kwargs)
        return densenet_retinanet(*args, backbone=self.backbone, **kwargs)

    def preprocess_image(self, inputs):
        return inputs / 127.5 - 1.0


def backbone(backbone_name) -> Backbone:
    """models/__init__.py:49-63."""
    if not any(k in backbone_name for k in ("resnet", "mobilenet", "densenet")):
        raise NotImplementedError("Backbone class for  '{}' not implemented.".format(backbone_name))
    return Backbone(backbone_name)


class FeatureExtractor:
    """retinanet.py:266-307.  `FeatureExtractor(retinanet_weight_path)(inp)` -> [f(P3), ..., f(P7)] NHWC float32.

    `retinanet_weight_path`: the reference loads a Keras .h5 by topology (retinanet.py:277-278); no HDF5 reader
    exists in this environment, so an `.npz` with the Appendix-B keys is accepted instead.
    """

    def __init__(self, retinanet_weight_path: Optional[str] = None, *, backbone: str = "mobilenet224_1.0",
                 weights: Optional[Dict[str, np.ndarray]] = None, precision: str = "bf16", device: int = 0, _cache=None):
        self.retinanet_model = backbone_spec_for(backbone)
        self._cache = _cache
        if _cache is None:
            from .transformer import _EngineCache
            if weights is None:
                weights = init_weights(backbone, vocab=8, num_layers=1)
            if retinanet_weight_path is not None:
                weights = dict(weights)
                weights.update(load_weights(retinanet_weight_path))
            vocab = int(weights["transformer/final_layer/kernel"].shape[1])
            layers = 1 + max(int(k.split("/")[3]) for k in weights if k.startswith("transformer/decoder/dec_layers/"))
            self._cache = _EngineCache(weights, backbone, layers, C.d_model, C.num_heads,
                                       int(weights["transformer/decoder/dec_layers/0/ffn1/kernel"].shape[1]), vocab, 4,
                                       precision, "log", device, C.START_ID, C.END_ID, False)

    def __call__(self, inp) -> List[torch.Tensor]:
        inp = torch.as_tensor(inp) if not isinstance(inp, torch.Tensor) else inp
        return self._cache.get(int(inp.shape[0]), 1).features(inp)

    call = __call__


def backbone_spec_for(name: str) -> RetinaNetSpec:
    return backbone(name).retinanet(C.NUM_OF_CLASSES)
