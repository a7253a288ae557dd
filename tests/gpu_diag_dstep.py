"""Developer diagnostic (not a pytest file): fused decoder (dstep_kernel) vs the per-operator chain and the oracle.
    python tests/gpu_diag_dstep.py
"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "fpn-mt-image-captioning_b200"), os.path.join(ROOT, "oracle")]
import fpnmt_oracle as O
from fpnmt.engine import Engine

def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))

def main():
    bb, B, S, L, V, T, N = "mobilenet224_1.0", 8, 256, 2, 1000, 12, 8
    w = O.caption_weights(bb, vocab=V, layers=L, seed=0, end_bias=6.0)
    Wv = O.W(w)
    img = O.test_images(B, S, seed=41)
    with torch.no_grad():
        mem = O.encoder(img, Wv, bb, num_layers=L, input_vocab_size=(S // 16) ** 2)
    gtok = torch.randint(4, V, (B, T), generator=torch.Generator().manual_seed(2)); gtok[:, 0] = 2
    ref_lg, _ = O.transformer_logits(mem, gtok, Wv, O.create_look_ahead_mask(T), T, num_layers=L)
    ids_ref, len_ref = O.predict_batch_cached(mem, Wv, T, N, 2, 3, num_layers=L, early_stop=False)
    res = {}
    for name, kw in (("chain_x3", dict(precision="bf16x3", decode_path="chain")), ("chain", dict(precision="bf16", decode_path="chain")),
                     ("fused", dict(precision="bf16", decode_path="fused", opts=("dstep_taps",)))):
        eng = Engine(w, backbone=bb, batch=B, beam=N, vocab=V, max_len=T, num_layers=L, image_size=S, use_graphs=False, **kw)
        for tt in (1, 2, T):
            lg = eng.decode_logits(mem.cuda(), gtok[:, :tt].int().cuda()).cpu()
            torch.cuda.synchronize()
            taps = {}
            for l in range(L):
                for k in (1, 2, 3):
                    x = eng.tap("dec%d_out%d" % (l, k)).cpu()
                    taps[(l, k)] = x.reshape(-1, 512)[:B * N]
            res[(name, tt)] = (lg, taps)
        t0 = time.time()
        ids, lens = eng.decode(early_stop=False)
        torch.cuda.synchronize()
        res[(name, "ids")] = (ids.numpy().copy(), lens.numpy().copy(), time.time() - t0)
        ids2, lens2 = eng.decode(early_stop=True)
        res[(name, "ids_es")] = (ids2.numpy().copy(), lens2.numpy().copy())
        eng.close()
    for tt in (1, 2, T):
        print("---- teacher forcing, %d step(s)" % tt)
        for name in ("chain", "fused"):
            lg, taps = res[(name, tt)]
            rlg, rtaps = res[("chain_x3", tt)]
            lp, lpr = torch.log_softmax(lg[..., :V], -1), torch.log_softmax(ref_lg[:, :tt], -1)
            print("%-6s logits rel vs oracle %.3e  vs chain_x3 %.3e  log-prob max err %.3e  argmax agree %.4f  finite %s" % (
                name, rel(lg, ref_lg[:, :tt]), rel(lg, rlg), float((lp - lpr).abs().max()),
                float((lp.argmax(-1) == lpr.argmax(-1)).float().mean()), bool(torch.isfinite(lg).all())))
            print("       taps vs chain_x3: " + "  ".join("L%d.out%d %.2e" % (l, k, rel(taps[(l, k)], rtaps[(l, k)])) for l in range(L) for k in (1, 2, 3)))
    for name in ("chain_x3", "chain", "fused"):
        ids, lens, dt = res[(name, "ids")]
        ids2, lens2 = res[(name, "ids_es")]
        print("%-8s generate: identical to oracle %d/%d images (fixed length), early-stop lens %s, %.1f ms" % (
            name, int(((ids == ids_ref).all(axis=1)).sum()), B, lens2.tolist(), dt * 1e3))
        if name == "fused":
            print("   fused ids[0]", ids[0].tolist()); print("   oracle      ", ids_ref[0].tolist())

if __name__ == "__main__":
    main()
