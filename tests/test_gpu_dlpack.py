"""The DLPack-typed C ABI (include/fpnmt_dlpack.h), Decoder.call's hidden states and the in-library NCCL communicator.

* An engine built and driven ONLY through raw ctypes + DLPack capsules (no fpnmt.engine): fpnmt_set_weight_dl from numpy
  arrays, fpnmt_generate_dl / fpnmt_encode_dl / fpnmt_decode_logits_dl on torch tensors - bit-identical to the raw-pointer
  entry points, which the other GPU tests hold to the oracle.
* Argument checking: wrong dtype / shape / device / non-contiguous tensors are FPNMT_ERR_INVALID naming the argument.
* fpnmt_decode_hidden == the oracle's Decoder.call restatement (/root/reference/models/transformer.py:321-341) within the
  BF16X3 stage tolerance (3e-3 rel-L2), and final_layer(hidden) reproduces fpnmt_decode_logits.
* fpnmt_comm_* with world = 1: the all-gather is the identity (the N > 1 exchange runs in bench.py under torchrun)."""
import ctypes as C

import numpy as np
import pytest
import torch

import fpnmt_oracle as O

pytestmark = pytest.mark.gpu

BB, S, L, V, T, N, B = "mobilenet224_1.0", 256, 2, 512, 8, 4, 2


def _raw_engine(w, precision=1):
    from fpnmt import _lib
    lib = _lib.load()
    c = _lib.FpnmtConfig()
    c.backbone, c.image_size, c.batch, c.beam, c.vocab, c.max_len = 0, S, B, N, V, T
    c.num_layers, c.d_model, c.num_heads, c.dff, c.precision, c.score_mode = L, 512, 8, 2048, precision, 0
    c.start_id, c.end_id, c.use_graphs = 2, 3, 1
    h = C.c_void_p()
    _lib.check(lib.fpnmt_create(C.byref(c), 0, C.byref(h)))
    for key, arr in w.items():
        a = np.ascontiguousarray(arr, dtype=np.float32)
        p, keep = _lib.dl_tensor(a)
        _lib.check(lib.fpnmt_set_weight_dl(h, key.encode(), p))
    _lib.check(lib.fpnmt_finalize_weights(h))
    return lib, h


def test_dlpack_entry_points_equal_raw_pointer_entry_points():
    from fpnmt import _lib
    from fpnmt.engine import Engine
    w = O.caption_weights(BB, vocab=V, layers=L, seed=0, end_bias=4.0)
    img = O.test_images(B, S, seed=7).contiguous()
    eng = Engine(w, backbone=BB, batch=B, beam=N, vocab=V, max_len=T, num_layers=L, image_size=S, precision="bf16x3")
    want_ids, want_len = eng.generate(img.cuda(), early_stop=True)
    want_mem = eng.encode(img.cuda()).clone()
    tok = torch.randint(4, V, (B, T), generator=torch.Generator().manual_seed(1)).int().cuda()
    tok[:, 0] = 2
    want_lg = eng.decode_logits(want_mem, tok).clone()
    want_feat = [f.clone() for f in eng.features(img.cuda())]
    eng.close()

    lib, h = _raw_engine(w)
    s = torch.cuda.current_stream().cuda_stream
    alive = []

    def dl(x):        # the capsule owns the DLTensor: keep every one alive until the end of the test
        p, keep = _lib.dl_tensor(x)
        alive.append(keep)
        return p, keep
    for src in (img.cuda(), img.pin_memory(), img.numpy()):            # device, pinned host, pageable numpy
        ids = torch.zeros((B, T), dtype=torch.int32, device="cuda")
        lens = torch.zeros((B,), dtype=torch.int32, device="cuda")
        (pi, k0), (po, k1), (pl, k2) = dl(src), dl(ids), dl(lens)
        _lib.check(lib.fpnmt_generate_dl(h, pi, po, pl, 1, None, s))
        torch.cuda.synchronize()
        assert torch.equal(ids.cpu(), want_ids) and torch.equal(lens.cpu(), want_len)
    ids_h, lens_h = np.zeros((B, T), np.int32), np.zeros((B,), np.int32)     # host outputs: the call synchronises
    (pi, k0), (po, k1), (pl, k2) = dl(img.cuda()), dl(ids_h), dl(lens_h)
    _lib.check(lib.fpnmt_generate_dl(h, pi, po, pl, 1, None, s))
    assert np.array_equal(ids_h, want_ids.numpy()) and np.array_equal(lens_h, want_len.numpy())
    mem = torch.zeros((B, 4, 512), device="cuda")
    (pi, k0), (pm, k1) = dl(img.cuda()), dl(mem)
    _lib.check(lib.fpnmt_encode_dl(h, pi, pm, s))
    lg = torch.zeros((B, T, V), device="cuda")
    (pt, k2), (pg, k3) = dl(tok), dl(lg)
    _lib.check(lib.fpnmt_decode_logits_dl(h, pm, pt, pg, s))
    feats = [torch.zeros((B, (S // 16) >> i, (S // 16) >> i, 512), device="cuda") for i in range(5)]
    fp = [dl(f) for f in feats]
    arr = (C.POINTER(_lib.DLTensor) * 5)(*[p for p, _ in fp])
    _lib.check(lib.fpnmt_features_dl(h, pi, arr, s))
    torch.cuda.synchronize()
    assert torch.equal(mem, want_mem)
    assert torch.equal(lg, want_lg), float((lg - want_lg).abs().max())
    for f, wf in zip(feats, want_feat):
        assert torch.equal(f, wf)

    # ---- argument checking: every rejected call names the argument and leaves the engine usable
    def rejected(fn, *a):
        rc = fn(*a)
        msg = lib.fpnmt_last_error().decode()
        assert rc == _lib.ERR_INVALID, (rc, msg)
        return msg
    ids = torch.zeros((B, T), dtype=torch.int32, device="cuda")
    lens = torch.zeros((B,), dtype=torch.int32, device="cuda")
    (po, k1), (pl, k2) = dl(ids), dl(lens)
    assert "images" in rejected(lib.fpnmt_generate_dl, h, dl(img.cuda().half())[0], po, pl, 1, None, s)              # dtype
    assert "images" in rejected(lib.fpnmt_generate_dl, h, dl(img.cuda()[:, :, :128].contiguous())[0], po, pl, 1, None, s)   # shape
    assert "row-major" in rejected(lib.fpnmt_generate_dl, h, dl(img.cuda().permute(0, 2, 1, 3))[0], po, pl, 1, None, s)     # strides
    assert "out_ids" in rejected(lib.fpnmt_generate_dl, h, dl(img.cuda())[0], dl(ids.float())[0], pl, 1, None, s)
    assert "both" in rejected(lib.fpnmt_generate_dl, h, dl(img.cuda())[0], po, dl(lens_h)[0], 1, None, s)
    assert "CUDA tensor" in rejected(lib.fpnmt_encode_dl, h, dl(img.cuda())[0], dl(mem.cpu())[0], s)
    assert "host tensor" in rejected(lib.fpnmt_set_weight_dl, h, b"x", dl(torch.zeros(3, device="cuda"))[0])
    _lib.check(lib.fpnmt_generate_dl(h, dl(img.cuda())[0], po, pl, 1, None, s))
    torch.cuda.synchronize()
    assert torch.equal(ids.cpu(), want_ids)
    lib.fpnmt_destroy(h)


def test_decoder_hidden_states_match_the_oracle_decoder():
    from fpnmt.transformer import Transformer
    w = O.caption_weights(BB, vocab=V, layers=L, seed=0)
    img = O.test_images(B, 512, seed=9).contiguous()       # the mirror keeps the reference's IMAGE_INPUT_SIZE
    Wv = O.W(w)
    tr = Transformer(L, 512, 8, 2048, 256, V, max_seq_len=T, backbone=BB, weights=w, precision="bf16x3")
    mem = tr.encoder(img.cuda(), False, None)
    tok = torch.randint(4, V, (B, T), generator=torch.Generator().manual_seed(3))
    tok[:, 0] = 2
    hid, attn = tr.decoder(tok, mem, False, None, None)                 # the reference's call shape (transformer.py:321)
    assert attn is None and tuple(hid.shape) == (B, T, 512)
    mask = O.create_look_ahead_mask(T, torch.float32)
    ref_hid, _ = O.decoder(tok, mem.cpu(), Wv, mask, T, L, 8)
    rel = float((hid.cpu() - ref_hid).norm() / ref_hid.norm())
    assert rel < 3e-3, rel
    logits, _ = tr(mem, tok, False, None)
    k, b = torch.from_numpy(w["transformer/final_layer/kernel"]), torch.from_numpy(w["transformer/final_layer/bias"])
    rel2 = float(((hid.cpu() @ k + b) - logits.cpu()).norm() / logits.cpu().norm())
    assert rel2 < 1e-3, rel2


def test_in_library_communicator_world_1():
    from fpnmt import _lib
    from fpnmt.dist import Communicator, allgather_captions
    comm = Communicator(1, 0, 0)
    ids = torch.arange(24, dtype=torch.int32, device="cuda").reshape(3, 8)
    lens = torch.tensor([1, 2, 3], dtype=torch.int32, device="cuda")
    a, l = comm.allgather(ids, lens)
    torch.cuda.synchronize()
    assert torch.equal(a, ids) and torch.equal(l, lens)
    a2, l2 = allgather_captions(ids, lens, 1, comm=comm)
    assert a2 is ids
    lib = _lib.load()
    out = torch.zeros((3, 8), dtype=torch.int32, device="cuda")
    ol = torch.zeros((3,), dtype=torch.int32, device="cuda")
    alive = []

    def dl(x):
        p, keep = _lib.dl_tensor(x)
        alive.append(keep)
        return p, keep
    _lib.check(lib.fpnmt_allgather_ids_dl(comm._c, dl(ids)[0], dl(lens)[0], dl(out)[0], dl(ol)[0], torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(out, ids) and torch.equal(ol, lens)
    rc = lib.fpnmt_allgather_ids_dl(comm._c, dl(ids)[0], dl(lens)[0], dl(out[:2].contiguous())[0], dl(ol)[0], None)
    assert rc == _lib.ERR_INVALID and "all_ids" in lib.fpnmt_last_error().decode()
    comm.close()
