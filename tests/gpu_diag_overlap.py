"""Diagnostic (not a test): does encode of batch i+1 overlap decode of batch i when two engines run on two streams of one GPU?
Timing only."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fpn-mt-image-captioning_b200"))
from fpnmt.engine import Engine            # noqa: E402
from fpnmt.weights import init_weights     # noqa: E402

B, N, V, T = 64, 8, 10000, 64
dev = torch.device("cuda", 0)
w = init_weights("resnet50", vocab=V, seed=0)
OPTS = tuple(o for o in os.environ.get("DIAG_OPTS", "").split(",") if o)
LANES = [int(x) for x in os.environ.get("DIAG_LANES", "1,2,3,4").split(",")]
g = torch.Generator().manual_seed(1234)
imgs = [(torch.rand(B, 512, 512, 3, generator=g) * 2 - 1).to(dev) for _ in range(2)]
print("opts", OPTS)
for L in LANES:
    eng = Engine(w, backbone="resnet50", batch=B, beam=N, vocab=V, max_len=T, precision="bf16", score_mode="log", device=0, opts=OPTS, lanes=L)

    def run(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in eng.generate_stream((imgs[i % 2] for i in range(n)), early_stop=False, to_host=False):
            pass
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    run(8)
    print("%d lanes: %.2f ms per batch" % (L, run(32)), flush=True)
    eng.close()
