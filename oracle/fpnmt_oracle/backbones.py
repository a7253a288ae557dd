"""Backbone CNNs restated from their upstream definitions.  TEST INFRASTRUCTURE ONLY.

None of these has source under /root/reference: the reference calls
`tf.keras.applications.mobilenet_v2.MobileNetV2` (models/mobilenet.py:61),
`keras_resnet.models.ResNet50(..., freeze_bn=True)` (models/resnet.py:99) and
`keras.applications.densenet.DenseNet121` (models/densenet.py:27,90) and taps the layers
named at mobilenet.py:64, resnet.py:112 (`outputs[1:]`) and densenet.py:93-96.
PARITY UNPINNED for this file: no reference test or fixture pins these numerics.

All functions take an NCHW tensor and the weight view `W`, and return (C3, C4, C5) in NCHW.
"""
from __future__ import annotations

import torch

from .ops import (W, avg_pool2, batch_norm, conv2d, depthwise_conv2d, max_pool, relu6)

__all__ = ["mobilenet_v2", "resnet50", "densenet121", "backbone_forward", "RN"]

RN = "transformer/encoder/feature_extractor/retinanet_model"

_MBV2 = [(1, 24, 2), (2, 24, 1), (3, 32, 2), (4, 32, 1), (5, 32, 1), (6, 64, 2), (7, 64, 1), (8, 64, 1),
         (9, 64, 1), (10, 96, 1), (11, 96, 1), (12, 96, 1), (13, 160, 2), (14, 160, 1), (15, 160, 1), (16, 320, 1)]


def mobilenet_v2(x: torch.Tensor, w: W):
    """Keras MobileNetV2 alpha=1.0, include_top=False.  BN eps 1e-3, ReLU6, no conv bias.
    Stride-2 convs: ZeroPadding2D(((0,1),(0,1))) + 'valid' (== TF SAME for even sizes)."""
    eps = 1e-3
    pad_s2 = ((0, 1), (0, 1))
    x = conv2d(x, w(RN + "/Conv1/kernel"), None, 2, pad_s2)
    x = relu6(batch_norm(x, w, RN + "/bn_Conv1", eps))
    # expanded_conv (block 0): no expansion
    x = depthwise_conv2d(x, w(RN + "/expanded_conv_depthwise/depthwise_kernel"), 1, "same")
    x = relu6(batch_norm(x, w, RN + "/expanded_conv_depthwise_BN", eps))
    x = conv2d(x, w(RN + "/expanded_conv_project/kernel"), None, 1, "same")
    x = batch_norm(x, w, RN + "/expanded_conv_project_BN", eps)
    taps = {}
    for k, cout, stride in _MBV2:
        p = RN + "/block_%d" % k
        inp = x
        cin = x.shape[1]
        x = conv2d(x, w(p + "_expand/kernel"), None, 1, "same")
        x = relu6(batch_norm(x, w, p + "_expand_BN", eps))
        if stride == 2:
            x = depthwise_conv2d(x, w(p + "_depthwise/depthwise_kernel"), 2, pad_s2)
        else:
            x = depthwise_conv2d(x, w(p + "_depthwise/depthwise_kernel"), 1, "same")
        x = relu6(batch_norm(x, w, p + "_depthwise_BN", eps))
        x = conv2d(x, w(p + "_project/kernel"), None, 1, "same")
        x = batch_norm(x, w, p + "_project_BN", eps)
        if stride == 1 and cin == cout:
            x = inp + x                                   # block_k_add
        taps[k] = x
    c3, c4 = taps[5], taps[12]                            # block_5_add, block_12_add
    x = conv2d(x, w(RN + "/Conv_1/kernel"), None, 1, "same")
    c5 = relu6(batch_norm(x, w, RN + "/Conv_1_bn", eps))  # out_relu
    return c3, c4, c5


def resnet50(x: torch.Tensor, w: W):
    """keras_resnet ResNet50 (v1, stride on the first 1x1 of a stage), BN eps 1e-5 frozen."""
    eps = 1e-5
    x = conv2d(x, w(RN + "/conv1/kernel"), None, 2, ((3, 3), (3, 3)))
    x = torch.relu(batch_norm(x, w, RN + "/bn_conv1", eps))
    x = max_pool(x, 3, 2, "same")                         # pool1
    outs = []
    for st, nblk in enumerate([3, 4, 6, 3]):
        for b in range(nblk):
            nm = "%d%s" % (st + 2, chr(ord("a") + b))
            stride = 2 if (b == 0 and st > 0) else 1
            y = conv2d(x, w(RN + "/res%s_branch2a/kernel" % nm), None, stride, "valid")
            y = torch.relu(batch_norm(y, w, RN + "/bn%s_branch2a" % nm, eps))
            y = conv2d(y, w(RN + "/res%s_branch2b/kernel" % nm), None, 1, ((1, 1), (1, 1)))
            y = torch.relu(batch_norm(y, w, RN + "/bn%s_branch2b" % nm, eps))
            y = conv2d(y, w(RN + "/res%s_branch2c/kernel" % nm), None, 1, "valid")
            y = batch_norm(y, w, RN + "/bn%s_branch2c" % nm, eps)
            if b == 0:
                sc = conv2d(x, w(RN + "/res%s_branch1/kernel" % nm), None, stride, "valid")
                sc = batch_norm(sc, w, RN + "/bn%s_branch1" % nm, eps)
            else:
                sc = x
            x = torch.relu(y + sc)
        outs.append(x)
    return outs[1], outs[2], outs[3]                      # outputs[1:] (resnet.py:112)


def densenet121(x: torch.Tensor, w: W):
    """keras.applications DenseNet121, taps conv{3,4,5}_block{12,24,16}_concat (pre final bn)."""
    eps = 1.001e-5
    x = conv2d(x, w(RN + "/conv1/conv/kernel"), None, 2, ((3, 3), (3, 3)))
    x = torch.relu(batch_norm(x, w, RN + "/conv1/bn", eps))
    x = max_pool(x, 3, 2, ((1, 1), (1, 1)))               # ZeroPadding2D(1) + 3x3 s2 valid (post-ReLU, so 0-pad is exact)
    outs = []
    for si, nblk in enumerate([6, 12, 24, 16]):
        stage = si + 2
        for b in range(1, nblk + 1):
            p = RN + "/conv%d_block%d" % (stage, b)
            y = torch.relu(batch_norm(x, w, p + "_0_bn", eps))
            y = conv2d(y, w(p + "_1_conv/kernel"), None, 1, "same")
            y = torch.relu(batch_norm(y, w, p + "_1_bn", eps))
            y = conv2d(y, w(p + "_2_conv/kernel"), None, 1, "same")
            x = torch.cat([x, y], dim=1)                  # conv{s}_block{b}_concat
        outs.append(x)
        if si < 3:
            p = RN + "/pool%d" % stage
            x = torch.relu(batch_norm(x, w, p + "_bn", eps))
            x = conv2d(x, w(p + "_conv/kernel"), None, 1, "same")
            x = avg_pool2(x)
    return outs[1], outs[2], outs[3]


def backbone_forward(name: str, x: torch.Tensor, w: W):
    if name.startswith("mobilenet"):
        return mobilenet_v2(x, w)
    if name == "resnet50":
        return resnet50(x, w)
    if name == "densenet121":
        return densenet121(x, w)
    raise ValueError("Backbone ('%s') is invalid." % name)
