"""Encode-only driver for ncu: ResNet-50 encoder of config C2 (B=64, 512x512), eager launches, 2 passes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fpn-mt-image-captioning_b200"))
import torch
from fpnmt.engine import Engine
from fpnmt.weights import init_weights

w = init_weights("resnet50", vocab=1000, seed=0)
eng = Engine(w, backbone="resnet50", batch=64, beam=8, vocab=1000, max_len=4, use_graphs=False)
img = torch.rand(64, 512, 512, 3, generator=torch.Generator().manual_seed(0)).cuda() * 2 - 1
for _ in range(2):
    mem = eng.encode(img)
torch.cuda.synchronize()
print("ok", float(mem.float().abs().mean()))
