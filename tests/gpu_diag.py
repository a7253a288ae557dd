"""GPU bring-up diagnostics (not a pytest): runs each check in-process, prints one line per check and writes
gpurun_out/diag.json.  Usage: python tests/gpu_diag.py [conv|beam|engine|gen|all] ...
"""
from __future__ import annotations

import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fpn-mt-image-captioning_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np
import torch
import torch.nn.functional as F

RESULTS = []


def record(name, ok, **kw):
    RESULTS.append(dict(name=name, ok=bool(ok), **kw))
    print(("PASS " if ok else "FAIL ") + name + " " + json.dumps(kw, default=str), flush=True)


def relerr(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max())


def conv_checks():
    from fpnmt.engine import conv2d
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(0)
    cases = [
        # name, N,H,W,Cin,Cout,kh,kw,pad, act, res_mode, bias
        ("dense128x64x64", 1, 1, 128, 64, 64, 1, 1, 0, 0, 0, False),
        ("dense512x512x512_bias", 1, 1, 512, 512, 512, 1, 1, 0, 0, 0, True),
        ("dense_ragged_rows37", 1, 1, 37, 128, 96, 1, 1, 0, 1, 0, True),
        ("dense_cout1000", 1, 1, 300, 512, 1000, 1, 1, 0, 0, 0, True),
        ("conv1x1_16x16", 2, 16, 16, 64, 256, 1, 1, 0, 0, 0, True),
        ("conv3x3_16x16", 2, 16, 16, 64, 64, 3, 3, 1, 1, 0, True),
        ("conv3x3_32x32_c256", 2, 32, 32, 256, 256, 3, 3, 1, 2, 0, True),
        ("conv3x3_8x8", 3, 8, 8, 128, 256, 3, 3, 1, 0, 0, True),
        ("conv3x3_4x4", 5, 4, 4, 64, 32, 3, 3, 1, 0, 0, True),
        ("conv3x3_2x2", 3, 2, 2, 64, 512, 3, 3, 1, 2, 0, True),
        ("conv3x3_1x1", 2, 1, 1, 64, 512, 3, 3, 1, 2, 0, True),
        ("conv3x3_cout1", 2, 16, 16, 256, 1, 3, 3, 1, 0, 0, True),
        ("conv1x1_cin24_cout144", 2, 16, 16, 24, 144, 1, 1, 0, 3, 0, True),
        ("conv1x1_res_same", 2, 16, 16, 128, 256, 1, 1, 0, 1, 1, True),
        ("conv1x1_res_up2", 2, 16, 16, 128, 256, 1, 1, 0, 0, 2, True),
        ("conv3x3_24x24_ragged", 1, 24, 24, 64, 64, 3, 3, 1, 0, 0, False),
        ("conv3x3_64x64_c256_big", 2, 64, 64, 256, 256, 3, 3, 1, 1, 0, True),
    ]
    for prec, tol in (("bf16", 2e-2), ("bf16x3", 2e-4)):
        for (name, N, H, W, Cin, Cout, kh, kw, pad, act, rm, hb) in cases:
            try:
                x = torch.randn(N, H, W, Cin, generator=g)
                k = (torch.randn(kh, kw, Cin, Cout, generator=g) / np.sqrt(kh * kw * Cin)).numpy()
                b = torch.randn(Cout, generator=g).numpy() * 0.5 if hb else None
                res = None
                if rm == 1:
                    res = torch.randn(N, H, W, Cout, generator=g)
                elif rm == 2:
                    res = torch.randn(N, H // 2, W // 2, Cout, generator=g)
                xd = x.cuda()
                y = conv2d(xd, k, b, act, (pad, pad), None if res is None else res.cuda(), rm, precision=prec)
                torch.cuda.synchronize()
                # fp64 reference on CPU
                xr = x.double().permute(0, 3, 1, 2)
                wr = torch.from_numpy(k).double().permute(3, 2, 0, 1)
                ref = F.conv2d(xr, wr, None if b is None else torch.from_numpy(b).double(), padding=pad)
                if rm == 1:
                    ref = ref + res.double().permute(0, 3, 1, 2)
                elif rm == 2:
                    ref = ref + res.double().permute(0, 3, 1, 2).repeat_interleave(2, 2).repeat_interleave(2, 3)
                if act == 1:
                    ref = torch.relu(ref)
                elif act == 2:
                    ref = torch.where(ref >= 0, ref, 0.2 * ref)
                elif act == 3:
                    ref = ref.clamp(0, 6)
                ref = ref.permute(0, 2, 3, 1)
                re_, mx = relerr(y.cpu(), ref)
                record("conv/%s/%s" % (prec, name), re_ < tol, rel=re_, maxabs=mx)
            except Exception as e:  # noqa
                record("conv/%s/%s" % (prec, name), False, err=repr(e)[:300])


def small_weights(backbone, vocab=512, layers=2, seed=0):
    from fpnmt.weights import init_weights
    gains = {"/model/conv2d_4": 48.0, "/model/conv2d_5": 48.0, "pyramid_regression": 20.0, "pyramid_classification": 20.0,
             "final_layer": 6.0}
    if backbone == "resnet50":
        gains.update({"_branch2c": 0.25, "C5_reduced": 0.01, "C4_reduced": 0.02, "C3_reduced": 0.1})
    return init_weights(backbone, vocab=vocab, seed=seed, num_layers=layers, randomize_bn=True, bias_std=0.02, gains=gains)


def engine_checks(backbones=("resnet50", "mobilenet224_1.0", "densenet121"), precs=("bf16x3", "bf16")):
    import fpnmt_oracle as O
    from fpnmt.engine import Engine
    B, S, L, V, T, N = 2, 256, 2, 512, 8, 4
    for bb in backbones:
        w = small_weights(bb, V, L)
        Wv = O.W(w)
        img = (torch.rand(B, S, S, 3, generator=torch.Generator().manual_seed(1)) * 2 - 1)
        taps = {}
        t0 = time.time()
        enc_ref = O.encoder(img, Wv, bb, num_layers=L, input_vocab_size=(S // 16) ** 2, taps=taps)
        print("oracle encoder %.1fs" % (time.time() - t0), flush=True)
        for prec in precs:
            tol = 3e-3 if prec == "bf16x3" else 8e-2
            try:
                eng = Engine(w, backbone=bb, batch=B, beam=N, vocab=V, max_len=T, num_layers=L, image_size=S,
                             precision=prec, use_graphs=False)
                mem = eng.encode(img.cuda())
                torch.cuda.synchronize()
                for nm in ("C3", "C4", "C5", "P3", "P4", "P5", "P6", "P7"):
                    got = eng.tap(nm).cpu().reshape(taps[nm].shape)
                    re_, mx = relerr(got, taps[nm])
                    record("engine/%s/%s/%s" % (bb, prec, nm), re_ < tol, rel=re_, maxabs=mx, std=float(taps[nm].std()))
                for i in range(5):
                    got = eng.tap("feat%d" % i).cpu().reshape(taps["features"][i].shape)
                    re_, mx = relerr(got, taps["features"][i])
                    record("engine/%s/%s/feat%d" % (bb, prec, i), re_ < tol, rel=re_, maxabs=mx, std=float(taps["features"][i].std()))
                for i in range(5):
                    got = eng.tap("tokens%d" % i).cpu().reshape(taps["tokens"][i].shape)
                    re_, mx = relerr(got, taps["tokens"][i])
                    record("engine/%s/%s/tokens%d" % (bb, prec, i), re_ < tol, rel=re_, maxabs=mx)
                for l in range(L):
                    got = eng.tap("enc_layer%d" % l).cpu().reshape(taps["enc_layer%d" % l].shape)
                    re_, mx = relerr(got, taps["enc_layer%d" % l])
                    record("engine/%s/%s/enc_layer%d" % (bb, prec, l), re_ < tol, rel=re_, maxabs=mx)
                re_, mx = relerr(mem.cpu(), enc_ref)
                record("engine/%s/%s/memory" % (bb, prec), re_ < tol, rel=re_, maxabs=mx)
                # teacher-forced logits
                gtok = torch.randint(4, V, (B, T), generator=torch.Generator().manual_seed(2))
                gtok[:, 0] = 2
                lg = eng.decode_logits(enc_ref.cuda(), gtok.int().cuda())
                torch.cuda.synchronize()
                mask = O.create_look_ahead_mask(T)
                ref_lg, _ = O.transformer_logits(enc_ref, gtok, Wv, mask, T, num_layers=L)
                lp = torch.log_softmax(lg.cpu(), -1)
                lpr = torch.log_softmax(ref_lg, -1)
                mx = float((lp - lpr).abs().max())
                record("engine/%s/%s/teacher_forced_logprob" % (bb, prec), mx < (2e-3 if prec == "bf16x3" else 1e-1), maxabs=mx,
                       argmax_agree=float((lp.argmax(-1) == lpr.argmax(-1)).float().mean()))
                # generate
                ids, lens = eng.generate(img.cuda(), early_stop=True)
                ref_ids, ref_len = O.predict_batch_cached(enc_ref, Wv, T, N, 2, 3, num_layers=L)
                same = bool((ids.numpy() == ref_ids).all() and (lens.numpy() == ref_len).all())
                record("engine/%s/%s/generate" % (bb, prec), same or prec == "bf16", ids=ids.numpy().tolist(), ref=ref_ids.tolist(),
                       lens=lens.numpy().tolist(), ref_len=ref_len.tolist())
                eng.close()
            except Exception as e:  # noqa
                traceback.print_exc()
                record("engine/%s/%s" % (bb, prec), False, err=repr(e)[:400])


def beam_checks():
    import fpnmt_oracle as O
    from fpnmt.engine import Engine
    from fpnmt.weights import init_weights
    B, N, V = 3, 4, 1000
    w = init_weights("mobilenet224_1.0", vocab=V, seed=0, num_layers=1)
    for mode in ("log", "prob"):
        try:
            eng = Engine(w, backbone="mobilenet224_1.0", batch=B, beam=N, vocab=V, max_len=4, num_layers=1, image_size=256,
                         score_mode=mode, use_graphs=False)
            g = torch.Generator().manual_seed(3)
            for trial in range(4):
                logits = torch.randn(B * N, V, generator=g) * (3.0 if trial % 2 else 0.3)
                if trial == 1:   # exact ties: identical rows (the reference's degenerate start)
                    logits = logits.reshape(B, N, V)[:, :1].repeat(1, N, 1).reshape(B * N, V)
                if trial == 2:   # ties inside a row
                    logits[:, 10] = logits[:, 500] = logits.max() + 1
                if mode == "log":
                    scores = -torch.rand(B * N, generator=g) * 3
                else:
                    scores = torch.rand(B * N, generator=g)
                if trial == 1:
                    scores = torch.zeros(B * N) if mode == "log" else torch.ones(B * N)
                if trial == 3 and mode == "prob":
                    scores = torch.zeros(B * N)   # underflowed products: reference picks flat indices 0..N-1
                par, tok, sc = eng.beam_step(logits.cuda(), scores.cuda())
                torch.cuda.synchronize()
                ok = True
                for b in range(B):
                    p, t, s = O.beam_step(logits[b * N:(b + 1) * N].numpy(), scores[b * N:(b + 1) * N].numpy(), mode)
                    ok &= bool((par[b * N:(b + 1) * N].cpu().numpy() == p).all() and (tok[b * N:(b + 1) * N].cpu().numpy() == t).all())
                    ok &= bool(np.allclose(sc[b * N:(b + 1) * N].cpu().numpy(), s, rtol=1e-5, atol=1e-6))
                record("beam/%s/trial%d" % (mode, trial), ok)
            eng.close()
        except Exception as e:  # noqa
            traceback.print_exc()
            record("beam/%s" % mode, False, err=repr(e)[:400])


if __name__ == "__main__":
    what = sys.argv[1:] or ["all"]
    print("device:", torch.cuda.get_device_name(0), flush=True)
    if "conv" in what or "all" in what:
        conv_checks()
    if "beam" in what or "all" in what:
        beam_checks()
    if "engine" in what or "all" in what:
        engine_checks()
    for a in what:
        if a.startswith("engine:"):
            engine_checks(backbones=(a.split(":")[1],))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag_%s.json" % "_".join(w.replace(":", "-") for w in what)), "w") as f:
        json.dump(RESULTS, f, indent=1, default=str)
    nfail = sum(1 for r in RESULTS if not r["ok"])
    print("SUMMARY: %d checks, %d failed" % (len(RESULTS), nfail))
