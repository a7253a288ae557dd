"""Developer diagnostic: in-kernel globaltimer timeline of one decode-chain op (needs `build.py --dbg-stamps`).
    FPNMT_DBG_OP=dec0_qkv python tests/gpu_diag_chain_timeline.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "fpn-mt-image-captioning_b200")]
from fpnmt import _lib
_lib.LIB_PATH = os.path.join(ROOT, "fpn-mt-image-captioning_b200", "libfpnmt_dbg.so")
from fpnmt.engine import Engine
from fpnmt.weights import init_weights
bb, B, N, V, T = "mobilenet224_1.0", 64, 8, 10000, 64
w = init_weights(bb, vocab=V, seed=0)
eng = Engine(w, backbone=bb, batch=B, beam=N, vocab=V, max_len=T, opts=tuple(o for o in os.environ.get("DIAG_OPTS", "").split(",") if o))
img = (torch.rand(B, 512, 512, 3) * 2 - 1).cuda()
eng.generate(img, early_stop=False)
torch.cuda.synchronize()
prof = eng.profile(iters=3)
print([(o["name"], round(o["us"], 2)) for o in prof["decode_step"][:7]])
