"""Small decode-only driver for ncu: MobileNetV2 encoder (cheap), C2-shaped decoder (B=64, beam 8, V=10000), T=12 steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fpn-mt-image-captioning_b200"))
import torch
from fpnmt.engine import Engine
from fpnmt.weights import init_weights

w = init_weights("mobilenet224_1.0", vocab=10000, seed=0)
eng = Engine(w, backbone="mobilenet224_1.0", batch=64, beam=8, vocab=10000, max_len=12, use_graphs=False,
             opts=tuple(o for o in os.environ.get("DIAG_OPTS", "").split(",") if o))
img = torch.rand(64, 512, 512, 3, generator=torch.Generator().manual_seed(0)).cuda() * 2 - 1
for _ in range(2):
    ids, lens = eng.generate(img, early_stop=False)
torch.cuda.synchronize()
print("ok", int(ids.sum()))
