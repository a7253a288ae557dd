"""Diagnostic: ancestry vs physical KV-cache mode on the bench's random-init weights (ids must be identical)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fpn-mt-image-captioning_b200"))
from fpnmt.engine import Engine
from fpnmt.weights import init_weights
bb = os.environ.get("DIAG_BB", "mobilenet224_1.0")
B, N, V, T = 64, 8, 10000, 64
w = init_weights(bb, vocab=V, seed=0)
g = torch.Generator().manual_seed(1234)
imgs = [(torch.rand(B, 512, 512, 3, generator=g) * 2 - 1).cuda() for _ in range(2)]
OPTS = tuple(o for o in os.environ.get("DIAG_OPTS", "").split(",") if o)
kw = dict(backbone=bb, batch=B, beam=N, vocab=V, max_len=T, precision="bf16", score_mode="log", opts=OPTS)
a = Engine(w, **kw)
a2 = Engine(w, **dict(kw, opts=OPTS + ("no_kv_share",)))
p = Engine(w, cache_mode="physical", **kw)
for i in range(2):
    ia, la = a.generate(imgs[i], early_stop=False, to_host=True)
    ib, lb = a2.generate(imgs[i], early_stop=False, to_host=True)
    ip, lp = p.generate(imgs[i], early_stop=False, to_host=True)
    print("batch", i, "a==a2", bool(torch.equal(ia, ib)), "a==p", bool(torch.equal(ia, ip)),
          "rows differing", int((ia != ip).any(dim=1).sum()), "first diff col", (ia != ip).any(dim=0).nonzero().flatten()[:4].tolist())
    print(" a[0,:12]", ia[0, :12].tolist(), "\n p[0,:12]", ip[0, :12].tolist())
for i in range(2):
    ia, _ = a.generate(imgs[1], early_stop=False, to_host=True)
    ip, _ = p.generate(imgs[1], early_stop=False, to_host=True)
    print("repeat", i, "a==p", bool(torch.equal(ia, ip)))
