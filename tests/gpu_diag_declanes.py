"""Diagnostic: decode-only and encode-only throughput with several engines on several streams of one GPU (timing only)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fpn-mt-image-captioning_b200"))
from fpnmt import _lib
if os.environ.get("DIAG_DBG"):
    _lib.LIB_PATH = os.path.join(ROOT, "fpn-mt-image-captioning_b200", "libfpnmt_dbg.so")
from fpnmt.engine import Engine
from fpnmt.weights import init_weights
B, N, V, T = 64, 8, 10000, 64
bb = os.environ.get("DIAG_BB", "mobilenet224_1.0")
OPTS = tuple(o for o in os.environ.get("DIAG_OPTS", "").split(",") if o)
NL = int(os.environ.get("DIAG_NL", "4"))
dev = torch.device("cuda", 0)
w = init_weights(bb, vocab=V, seed=0)
engs = [Engine(w, backbone=bb, batch=B, beam=N, vocab=V, max_len=T, precision="bf16", score_mode="log", device=0, opts=OPTS) for _ in range(NL)]
img = (torch.rand(B, 512, 512, 3) * 2 - 1).to(dev)
streams = [torch.cuda.Stream(dev) for _ in range(NL)]
for e in engs:
    e.generate(img, early_stop=False, to_host=False)
torch.cuda.synchronize()
def run(n, L, what):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        with torch.cuda.stream(streams[i % L]):
            if what == "dec":
                engs[i % L].decode(early_stop=False, to_host=False)
            else:
                engs[i % L].encode(img)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
for what in tuple(os.environ.get("DIAG_WHAT", "dec").split(",")):
    for L in sorted({1, 2, NL // 2, NL}):
        run(4, L, what)
        ms = run(16, L, what)
        print("%s-only, %d lanes: %.2f ms per batch%s" % (what, L, ms, " (%.1f us/step)" % (ms * 1e3 / T) if what == "dec" else ""), flush=True)
