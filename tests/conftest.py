import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "fpn-mt-image-captioning_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """libfpnmt.so, built in-tree (nvcc cross-compiles without a GPU)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fpnmt_build", os.path.join(ROOT, "fpn-mt-image-captioning_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=False, verbose=False)


TEST_GAINS = {"/model/conv2d_4": 48.0, "/model/conv2d_5": 48.0, "pyramid_regression": 20.0, "pyramid_classification": 20.0,
              "final_layer": 6.0}


def small_weights(backbone, vocab=512, layers=2, seed=0):
    """Random weights with non-trivial BN statistics / biases and gains that keep every stage's signal O(1)
    (at Keras-default init the co-attention's 1/(H*W) scale makes head outputs vanish)."""
    from fpnmt.weights import init_weights
    gains = dict(TEST_GAINS)
    if backbone == "resnet50":
        gains.update({"_branch2c": 0.25, "C5_reduced": 0.01, "C4_reduced": 0.02, "C3_reduced": 0.1})
    return init_weights(backbone, vocab=vocab, seed=seed, num_layers=layers, randomize_bn=True, bias_std=0.02, gains=gains)
