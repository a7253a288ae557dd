"""Finished-beam bookkeeping + length penalty (flagged EXTENSION; fpnmt_config.finished_beams / length_penalty).
With the extension off the engine is the reference (every other test); with it on, the engine must reproduce the oracle's
`predict_batch_cached(finished_beams=True, length_penalty=alpha)`: frozen beams compete with their final score, ranking by
score / ((5 + len) / 6)^alpha, an image stops when its best beam is frozen.  BF16X3 mode, 24 images with non-degenerate
captions, ~1/3 of them stopping early; >= 95 % identical sequences (the length-penalty table is computed with powf on the host
and numpy in the oracle: a 1-ulp difference may reorder an exact near-tie).  alpha = 0 with the reference's beam initialisation
must equal the reference itself."""
import numpy as np
import pytest
import torch

import fpnmt_oracle as O

pytestmark = pytest.mark.gpu

BB, S, L, V, T, N, B = "mobilenet224_1.0", 256, 2, 1000, 16, 8, 24


@pytest.fixture(scope="module")
def subject():
    w = O.caption_weights(BB, vocab=V, layers=L, seed=0, end_bias=6.0)
    Wv = O.W(w)
    img = O.test_images(B, S, seed=41)
    with torch.no_grad():
        mem = O.encoder(img, Wv, BB, num_layers=L, input_vocab_size=(S // 16) ** 2)
    return dict(w=w, Wv=Wv, img=img, mem=mem)


@pytest.mark.parametrize("true_beam,alpha", [(False, 0.0), (True, 0.0), (True, 0.8)])
def test_finished_beams_and_length_penalty_match_oracle(subject, true_beam, alpha):
    from fpnmt.engine import Engine
    s = subject
    ref_ids, ref_len = O.predict_batch_cached(s["mem"], s["Wv"], T, N, 2, 3, num_layers=L, true_beam=true_beam,
                                              finished_beams=True, length_penalty=alpha)
    eng = Engine(s["w"], backbone=BB, batch=B, beam=N, vocab=V, max_len=T, num_layers=L, image_size=S, precision="bf16x3",
                 true_beam=true_beam, finished_beams=True, length_penalty=alpha)
    eng.decode_logits(s["mem"].cuda(), torch.full((B, 1), 2, dtype=torch.int32).cuda())      # decode from the oracle's memory
    ids, lens = eng.decode(early_stop=True)
    ids2, lens2 = eng.decode(early_stop=False)
    eng.close()
    assert torch.equal(ids, ids2) and torch.equal(lens, lens2)
    same = (ids.numpy() == ref_ids).all(axis=1) & (lens.numpy() == ref_len)
    assert same.mean() >= 0.95, (same, lens.tolist(), ref_len.tolist())
    if not true_beam:    # alpha = 0 and identical beams: exactly the reference's own result
        base_ids, base_len = O.predict_batch_cached(s["mem"], s["Wv"], T, N, 2, 3, num_layers=L)
        assert (ref_ids == base_ids).all() and (ref_len == base_len).all()
    else:                # the extension is not a no-op on these weights
        plain_ids, plain_len = O.predict_batch_cached(s["mem"], s["Wv"], T, N, 2, 3, num_layers=L, true_beam=True)
        assert ((plain_ids != ref_ids).any(axis=1) | (plain_len != ref_len)).sum() >= 3
        assert (ref_len < T).sum() >= 4


def test_length_penalty_without_finished_beams_is_rejected():
    from fpnmt._lib import FpnmtError
    from fpnmt.engine import Engine
    w = O.caption_weights(BB, vocab=512, layers=L, seed=1)
    with pytest.raises(FpnmtError):
        Engine(w, backbone=BB, batch=2, beam=4, vocab=512, max_len=8, num_layers=L, image_size=S, length_penalty=0.6)
