"""fpnmt — B200-native batched caption inference for the FPN + Multi-Transformer captioner.

Host-side mirror of the reference's Python builder API over libfpnmt.so (hand-written sm_100a CUDA behind a C ABI).
"""
from . import config  # noqa: F401
from .weights import init_weights, load_weights, save_weights, model_spec  # noqa: F401

__all__ = ["config", "init_weights", "load_weights", "save_weights", "model_spec"]
