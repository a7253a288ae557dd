"""CPU oracle for the batched caption-inference hot path.  TEST INFRASTRUCTURE ONLY.

A plain PyTorch-CPU restatement (fp32, optional fp64) of the reference's forward:
/root/reference/utils/pipeline.py:82-154, models/transformer.py, models/retinanet.py,
models/coattention.py, layers/_misc.py and the un-vendored Keras backbones the reference
wires.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` leg may import this package; the product
(`fpn-mt-image-captioning_b200/`) never does.

PARITY PINNING.  The reference ships no golden vectors, known-answer tests or fixtures
for this path and TensorFlow/Keras cannot be installed here.  The
oracle is pinned as far as the container allows by `tests/golden/make_golden.py`, which
imports the reference's OWN Python modules from /root/reference under a small numpy
stand-in for the TensorFlow primitives they call and records their outputs
(positional encoding, look-ahead mask, scaled-dot-product attention, MultiHeadAttention,
EncoderLayer, DecoderLayer, Decoder, CoAttention_CNN and the `Pipeline.predict` beam loop);
`tests/test_oracle_golden.py` checks this oracle against those vectors.  The Keras
backbones/functional-model plumbing (MobileNetV2, keras_resnet, DenseNet) have no source
under /root/reference: for them parity is UNPINNED and the restatement here is the
definition (stated in DESIGN.md).
"""
from .ops import *            # noqa: F401,F403
from .backbones import *      # noqa: F401,F403
from .model import *          # noqa: F401,F403
from .decode import *         # noqa: F401,F403
from .testing import *        # noqa: F401,F403
