"""Diagnostic: repeatability of the streamed throughput (device-resident and pinned-host inputs) within one process."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fpn-mt-image-captioning_b200"))
from fpnmt.engine import Engine
from fpnmt.weights import init_weights
bb = os.environ.get("DIAG_BB", "resnet50")
NL = int(os.environ.get("DIAG_NL", "4"))
STEPS = int(os.environ.get("DIAG_STEPS", "12"))
B, N, V, T = 64, 8, 10000, 64
w = init_weights(bb, vocab=V, seed=0)
eng = Engine(w, backbone=bb, batch=B, beam=N, vocab=V, max_len=T, precision="bf16", score_mode="log", lanes=NL)
g = torch.Generator().manual_seed(1234)
host = [(torch.rand(B, 512, 512, 3, generator=g) * 2 - 1).pin_memory() for _ in range(2)]
dev = [h.cuda() for h in host]
def run(src, to_host):
    for _ in eng.generate_stream((src[i % 2] for i in range(STEPS)), early_stop=False, to_host=to_host):
        pass
for _ in range(2):
    run(dev, False)
torch.cuda.synchronize()
for name, src, th in (("device", dev, False), ("host", host, True), ("device", dev, False), ("host", host, True)):
    res = []
    for r in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(src, th)
        torch.cuda.synchronize()
        res.append(B * STEPS / (time.perf_counter() - t0))
    print(name, "lanes", NL, " ".join("%.0f" % x for x in res), flush=True)
