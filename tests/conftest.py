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


def small_weights(backbone, vocab=512, layers=2, seed=0):
    """Test weights: reference variable tree and distributions, gains + calibrated BatchNorm statistics so that every
    stage carries an O(1) input-dependent signal (see oracle/fpnmt_oracle/testing.py)."""
    import fpnmt_oracle as O
    return O.test_weights(backbone, vocab=vocab, layers=layers, seed=seed)
