"""Diagnostic: batch-permutation invariance of the encoder memory under kernel_opts (DIAG_OPTS)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "fpn-mt-image-captioning_b200"), os.path.join(ROOT, "oracle")]
import fpnmt_oracle as O
from fpnmt.engine import Engine
bb, Bf, Nf, Vf, Tf = "mobilenet224_1.0", 4, 8, 1000, 12
w = O.test_weights(bb, vocab=Vf, layers=2, seed=2)
imgs = O.test_images(Bf, 512, seed=9)
perm = torch.tensor([2, 0, 3, 1])
for opts in ((), ("no_pdl",), ("no_tma_store",)):
    eng = Engine(w, backbone=bb, batch=Bf, beam=Nf, vocab=Vf, max_len=Tf, num_layers=2, image_size=512, opts=opts, use_graphs=not os.environ.get("DIAG_EAGER"))
    m1 = eng.encode(imgs.cuda()).cpu()
    m1b = eng.encode(imgs.cuda()).cpu()
    m2 = eng.encode(imgs[perm].cuda()).cpu()
    bad = {}
    for nm in ("C3", "C4", "C5", "P3", "P4", "P5", "P6", "P7", "feat0", "feat4", "tokens0", "tokens4", "enc_layer0"):
        eng.encode(imgs.cuda()); a = eng.tap(nm).cpu().clone()
        eng.encode(imgs[perm].cuda()); b = eng.tap(nm).cpu()
        a = a.reshape(Bf, -1); b = b.reshape(Bf, -1)
        bad[nm] = float((a[perm] - b).abs().max())
    print(opts, "deterministic", bool(torch.equal(m1, m1b)), "perm-invariant", bool(torch.equal(m1[perm], m2)), {k: v for k, v in bad.items() if v > 0})
    eng.close()
