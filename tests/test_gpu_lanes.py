"""Lanes (fpnmt_submit / fpnmt_collect; fpnmt_config.lanes): several batches in flight on one GPU, the encoder of batch i+1 under
the decode of batch i.  Each lane is a complete engine running the same kernels in the same order as fpnmt_generate, so every
result must be BIT-IDENTICAL to the one-shot call on the same images - from host and from device memory, fixed-length and with
early stop, with more batches than lanes and with fewer - and agree with the oracle's Pipeline.predict restatement
(/root/reference/utils/pipeline.py:82-154)."""
import pytest
import torch

import fpnmt_oracle as O

pytestmark = pytest.mark.gpu

BB, S, L, V, T, N, B = "mobilenet224_1.0", 256, 2, 1000, 12, 8, 4


@pytest.mark.parametrize("lanes", [2, 3])
def test_lanes_stream_equals_one_shot_generate(lanes):
    from fpnmt.engine import Engine
    w = O.caption_weights(BB, vocab=V, layers=L, seed=0, end_bias=6.0)
    batches = [O.test_images(B, S, seed=50 + i) for i in range(5)]
    one = Engine(w, backbone=BB, batch=B, beam=N, vocab=V, max_len=T, num_layers=L, image_size=S, precision="bf16x3")
    want = {es: [tuple(t.clone() for t in one.generate(b.cuda(), early_stop=es)) for b in batches] for es in (False, True)}
    one.close()
    eng = Engine(w, backbone=BB, batch=B, beam=N, vocab=V, max_len=T, num_layers=L, image_size=S, precision="bf16x3", lanes=lanes)
    pinned = [b.pin_memory() for b in batches]
    on_dev = [b.cuda() for b in batches]
    for es in (False, True):
        for src in (pinned, on_dev, pinned[:1], pinned[:lanes]):
            got = list(eng.generate_stream(iter(src), early_stop=es))
            assert len(got) == len(src)
            for (ids, lens), (wi, wl) in zip(got, want[es]):
                assert torch.equal(ids, wi) and torch.equal(lens, wl)
        got_dev = list(eng.generate_stream(iter(on_dev), early_stop=es, to_host=False))
        for (ids, lens), (wi, wl) in zip(got_dev, want[es]):
            assert ids.is_cuda and torch.equal(ids.cpu(), wi) and torch.equal(lens.cpu(), wl)
    # lane 0 is also the engine of the single-batch entry points
    ids, lens = eng.generate(on_dev[2], early_stop=True)
    assert torch.equal(ids, want[True][2][0])
    assert eng.launch_count > 0
    eng.close()
    # ... and the captions are the oracle's (fp32-class mode: identical sequences)
    Wv = O.W(w)
    for b, (wi, wl) in zip(batches[:2], want[True][:2]):
        for j in range(B):
            ref = O.predict_reference(b[j], Wv, T, N, 2, 3, BB, num_layers=L, mode="log")
            assert wi[j, :int(wl[j])].tolist() == list(ref)


def test_lane_call_order_errors():
    from fpnmt._lib import FpnmtError
    from fpnmt.engine import Engine
    w = O.caption_weights(BB, vocab=512, layers=L, seed=1)
    eng = Engine(w, backbone=BB, batch=2, beam=4, vocab=512, max_len=8, num_layers=L, image_size=S, lanes=2)
    x = O.test_images(2, S, seed=3).cuda()
    with pytest.raises(FpnmtError):
        eng.collect(0)                        # nothing submitted
    with pytest.raises(FpnmtError):
        eng.submit(2, x)                      # lane out of range
    keep = eng.submit(1, x)
    with pytest.raises(FpnmtError):
        eng.submit(1, x)                      # lane busy
    ids, lens = eng.collect(1)
    assert tuple(ids.shape) == (2, 8)
    del keep
    eng.close()
    with pytest.raises(FpnmtError):
        Engine(w, backbone=BB, batch=2, beam=4, vocab=512, max_len=8, num_layers=L, image_size=S, lanes=17)        # the handle holds at most 16 lanes
