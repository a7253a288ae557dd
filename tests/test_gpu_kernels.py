"""GPU parity of single kernels through the C ABI, against the CPU oracle / fp64 arithmetic on the same seeded inputs.

tolerances: BF16X3 (the parity mode: 3-term split-bf16 products, fp32 accumulate) relative L2 error <= 2e-4;
BF16 (the fast mode: bf16 operands + bf16 activation storage) relative L2 error <= 2e-2.  Integer outputs
(beam parents / tokens) must be bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import fpnmt_oracle as O

pytestmark = pytest.mark.gpu

CONV_CASES = [
    # name, N,H,W,Cin,Cout,kh,kw,pad, act, res_mode, bias
    ("dense128x64x64", 1, 1, 128, 64, 64, 1, 1, 0, 0, 0, False),
    ("dense512x512x512_bias", 1, 1, 512, 512, 512, 1, 1, 0, 0, 0, True),
    ("dense_ragged_rows37", 1, 1, 37, 128, 96, 1, 1, 0, 1, 0, True),
    ("dense_cout1000", 1, 1, 300, 512, 1000, 1, 1, 0, 0, 0, True),
    ("dense_k2048", 1, 1, 256, 2048, 512, 1, 1, 0, 2, 1, True),
    ("conv1x1_16x16", 2, 16, 16, 64, 256, 1, 1, 0, 0, 0, True),
    ("conv3x3_16x16", 2, 16, 16, 64, 64, 3, 3, 1, 1, 0, True),
    ("conv3x3_32x32_c256", 2, 32, 32, 256, 256, 3, 3, 1, 2, 0, True),
    ("conv3x3_8x8", 3, 8, 8, 128, 256, 3, 3, 1, 0, 0, True),
    ("conv3x3_4x4", 5, 4, 4, 64, 32, 3, 3, 1, 0, 0, True),
    ("conv3x3_2x2", 3, 2, 2, 64, 512, 3, 3, 1, 2, 0, True),
    ("conv3x3_1x1", 2, 1, 1, 64, 512, 3, 3, 1, 2, 0, True),
    ("conv3x3_cout1", 2, 16, 16, 256, 1, 3, 3, 1, 0, 0, True),
    ("conv1x1_cin24_cout144", 2, 16, 16, 24, 144, 1, 1, 0, 3, 0, True),
    ("conv1x1_cin96_cout24", 1, 32, 32, 96, 24, 1, 1, 0, 0, 1, True),
    ("conv1x1_res_same", 2, 16, 16, 128, 256, 1, 1, 0, 1, 1, True),
    ("conv1x1_res_up2", 2, 16, 16, 128, 256, 1, 1, 0, 0, 2, True),
    ("conv3x3_24x24_ragged", 1, 24, 24, 64, 64, 3, 3, 1, 0, 0, False),
    ("conv3x3_64x64_c256", 2, 64, 64, 256, 256, 3, 3, 1, 1, 0, True),
    ("conv7x7_pad3", 1, 16, 16, 8, 32, 7, 7, 3, 0, 0, True),
]


def _conv_ref(x, k, b, pad, act, res, rm):
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
    return ref.permute(0, 2, 3, 1)


@pytest.mark.parametrize("prec,tol", [("bf16x3", 2e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_igemm_conv(case, prec, tol):
    from fpnmt.engine import conv2d
    name, N, H, W, Cin, Cout, kh, kw, pad, act, rm, hb = case
    g = torch.Generator().manual_seed(hash(name) % 1000)
    x = torch.randn(N, H, W, Cin, generator=g)
    k = (torch.randn(kh, kw, Cin, Cout, generator=g) / np.sqrt(kh * kw * Cin)).numpy()
    b = torch.randn(Cout, generator=g).numpy() * 0.5 if hb else None
    res = None
    if rm == 1:
        res = torch.randn(N, H, W, Cout, generator=g)
    elif rm == 2:
        res = torch.randn(N, H // 2, W // 2, Cout, generator=g)
    y = conv2d(x.cuda(), k, b, act, (pad, pad), None if res is None else res.cuda(), rm, precision=prec).cpu()
    ref = _conv_ref(x, k, b, pad, act, res, rm)
    rel = float((y.double() - ref).norm() / ref.norm())
    assert rel < tol, "%s %s rel %g" % (name, prec, rel)


@pytest.mark.parametrize("force_bn", [32, 64, 128, 256])
def test_igemm_all_tile_widths(force_bn):
    from fpnmt.engine import conv2d
    g = torch.Generator().manual_seed(force_bn)
    x = torch.randn(2, 16, 16, 128, generator=g)
    k = (torch.randn(3, 3, 128, 320, generator=g) / np.sqrt(9 * 128)).numpy()
    y = conv2d(x.cuda(), k, None, 0, (1, 1), precision="bf16x3", force_bn=force_bn).cpu()
    ref = _conv_ref(x, k, None, 1, 0, None, 0)
    assert float((y.double() - ref).norm() / ref.norm()) < 2e-4


def test_igemm_bulk_store_epilogue_many_tiles_per_cta_is_deterministic():
    """Persistent CTAs reuse their output staging buffers tile after tile while earlier bulk tensor stores may still be reading
    them.  144 filters in a 256-wide tile leave one epilogue warp with a single channel pair per tile - the case in which letting one
    bulk group stay in flight restaged a buffer under the TMA unit (found by test_batch_invariance_and_determinism_512).  Many
    tiles per CTA, result against the fp64 reference and bit-identical across runs; also with a residual (its prefetch lands in
    the same buffers) and in the 2-D pixel-tile geometry (3x3)."""
    from fpnmt.engine import conv2d
    g = torch.Generator().manual_seed(7)
    for (n, h, w, cin, cout, kk, rm) in ((2, 256, 256, 32, 144, 1, 0), (2, 128, 128, 64, 256, 1, 1), (2, 96, 96, 64, 192, 3, 0)):
        x = torch.randn(n, h, w, cin, generator=g)
        k = (torch.randn(kk, kk, cin, cout, generator=g) / np.sqrt(kk * kk * cin)).numpy()
        b = (torch.randn(cout, generator=g) * 0.5).numpy()
        res = torch.randn(n, h, w, cout, generator=g) if rm else None
        kw = dict(bias=b, act=1, pad=(kk // 2, kk // 2), residual=None if res is None else res.cuda(), res_mode=rm, precision="bf16")
        y0 = conv2d(x.cuda(), k, **kw).cpu()
        ref = _conv_ref(x, k, b, kk // 2, 1, res, rm)
        assert float((y0.double() - ref).norm() / ref.norm()) < 2e-2
        for _ in range(3):
            assert torch.equal(conv2d(x.cuda(), k, **kw).cpu(), y0)


def test_igemm_zero_input_and_linearity():
    from fpnmt.engine import conv2d
    g = torch.Generator().manual_seed(1)
    k = (torch.randn(3, 3, 64, 64, generator=g) / 24).numpy()
    x1, x2 = torch.randn(1, 16, 16, 64, generator=g), torch.randn(1, 16, 16, 64, generator=g)
    z = conv2d(torch.zeros(1, 16, 16, 64).cuda(), k, None, 0, (1, 1), precision="bf16x3")
    assert float(z.abs().max()) == 0.0
    y1 = conv2d(x1.cuda(), k, None, 0, (1, 1), precision="bf16x3")
    y2 = conv2d(x2.cuda(), k, None, 0, (1, 1), precision="bf16x3")
    y12 = conv2d((x1 + x2).cuda(), k, None, 0, (1, 1), precision="bf16x3")
    assert float((y12 - y1 - y2).abs().max()) < 2e-4 * float(y12.abs().max())


@pytest.fixture(scope="module")
def beam_engines():
    from fpnmt.engine import Engine
    from fpnmt.weights import init_weights
    B, N, V = 3, 4, 1000
    w = init_weights("mobilenet224_1.0", vocab=V, seed=0, num_layers=1)
    engs = {m: Engine(w, backbone="mobilenet224_1.0", batch=B, beam=N, vocab=V, max_len=4, num_layers=1, image_size=256,
                      score_mode=m, use_graphs=False) for m in ("log", "prob")}
    yield engs, B, N, V
    for e in engs.values():
        e.close()


@pytest.mark.parametrize("mode", ["log", "prob"])
@pytest.mark.parametrize("trial", ["random_flat", "identical_rows", "ties_in_row", "peaked", "underflow"])
def test_beam_step_bit_exact(beam_engines, mode, trial):
    engs, B, N, V = beam_engines
    eng = engs[mode]
    g = torch.Generator().manual_seed(len(trial) * 7 + (mode == "log"))
    logits = torch.randn(B * N, V, generator=g) * (3.0 if trial == "peaked" else 0.3)
    scores = -torch.rand(B * N, generator=g) * 3 if mode == "log" else torch.rand(B * N, generator=g)
    if trial == "identical_rows":      # the reference's start state (pipeline.py:101-102)
        logits = logits.reshape(B, N, V)[:, :1].repeat(1, N, 1).reshape(B * N, V)
        scores = torch.zeros(B * N) if mode == "log" else torch.ones(B * N)
    if trial == "ties_in_row":
        logits[:, 10] = logits[:, 500] = logits.max() + 1
    if trial == "underflow":
        if mode == "log":
            pytest.skip("underflow is a property of the probability-product score")
        scores = torch.zeros(B * N)
    par, tok, sc = eng.beam_step(logits.cuda(), scores.cuda())
    par, tok, sc = par.cpu().numpy(), tok.cpu().numpy(), sc.cpu().numpy()
    for b in range(B):
        sl = slice(b * N, (b + 1) * N)
        p, t, s = O.beam_step(logits[sl].numpy(), scores[sl].numpy(), mode)
        assert (par[sl] == p).all() and (tok[sl] == t).all(), (trial, b, par[sl], p, tok[sl], t)
        assert np.allclose(sc[sl], s, rtol=1e-5, atol=1e-6)
    if trial == "underflow":
        assert tok[:N].tolist() == [0, 1, 2, 3] and par[:N].tolist() == [0, 0, 0, 0]


# ---------------------------------------------------------------------------------------------------------------------
# skinny-row Dense kernel of the decoder step (tgemm): plain and cluster-LayerNorm epilogues
DENSE_CASES = [
    # name, R, K, F, act, residual, layernorm, force_bn
    ("o_proj_512", 512, 512, 512, 0, False, False, 0),
    ("qkv_1536_bn64", 512, 512, 1536, 0, False, False, 0),
    ("ffn1_leaky_2048", 512, 512, 2048, 2, False, False, 0),
    ("ffn2_k2048_res", 512, 2048, 512, 0, True, False, 0),
    ("vocab_1000_ragged_features", 192, 512, 1000, 0, False, False, 0),
    ("ragged_rows_100_bn32", 100, 512, 512, 1, True, False, 32),
    ("ragged_rows_100_bn64", 100, 512, 256, 0, False, False, 64),
    ("multi_rowtile_stationary", 1024, 512, 10000, 0, False, False, 0),
    ("ln_o_proj", 512, 512, 512, 0, True, True, 0),
    ("ln_ffn2_k2048", 512, 2048, 512, 0, True, True, 0),
    ("ln_ragged_rows_40", 40, 512, 512, 0, True, True, 0),
    ("ln_no_residual", 64, 512, 512, 0, False, True, 0),
    # wide-row kernel (tgemmw.cu, FPNMT_OPT_TGEMM_WIDE): 128 (64) rows per CTA, ring-streamed K, two CTAs per SM; bf16 only
    ("wide_qkv_1536", 512, 512, 1536, 0, False, False, 128),
    ("wide_ffn1_leaky_2048", 512, 512, 2048, 2, False, False, 128),
    ("wide_vocab_10000_multi_rowtile", 1024, 512, 10000, 0, False, False, 128),
    ("wide_vocab_1000_ragged_rows_192", 192, 512, 1000, 1, False, False, 128),
    ("wide_rows_40_bn64", 40, 512, 256, 0, False, False, 128),
    ("wide_ln_o_proj", 512, 512, 512, 0, True, True, 128),
    ("wide_ln_ffn2_k2048", 512, 2048, 512, 0, True, True, 128),
    ("wide_ln_ragged_rows_300", 300, 512, 512, 0, True, True, 128),
    ("wide_ln_rows_40_bn64", 40, 2048, 512, 0, True, True, 128),
]


@pytest.mark.parametrize("prec,tol", [("bf16x3", 2e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("case", DENSE_CASES, ids=[c[0] for c in DENSE_CASES])
def test_tgemm_dense(case, prec, tol):
    from fpnmt.engine import dense
    name, R, K, Fo, act, has_res, ln, fbn = case
    if fbn >= 128 and prec != "bf16":
        pytest.skip("the wide-row kernel is a bf16-mode kernel")
    g = torch.Generator().manual_seed(abs(hash(name)) % (2 ** 31))
    x = torch.randn(R, K, generator=g)
    k = (torch.randn(K, Fo, generator=g) / np.sqrt(K)).numpy()
    b = torch.randn(Fo, generator=g).numpy() * 0.5
    res = torch.randn(R, Fo, generator=g) if has_res else None
    gam = (torch.rand(Fo, generator=g) + 0.5).numpy() if ln else None
    bet = torch.randn(Fo, generator=g).numpy() if ln else None
    y = dense(x.cuda(), k, b, act, None if res is None else res.cuda(), gam, bet, 1e-6, prec, fbn).cpu()
    ref = x.double() @ torch.from_numpy(k).double() + torch.from_numpy(b).double()
    if has_res:
        ref = ref + res.double()
    if act == 1:
        ref = torch.relu(ref)
    elif act == 2:
        ref = torch.where(ref >= 0, ref, 0.2 * ref)
    if ln:
        mean = ref.mean(-1, keepdim=True)
        var = ((ref - mean) ** 2).mean(-1, keepdim=True)
        ref = (ref - mean) / torch.sqrt(var + 1e-6) * torch.from_numpy(gam).double() + torch.from_numpy(bet).double()
    err = float((y.double() - ref).norm() / ref.norm())
    assert err < tol, (name, prec, err)
    assert torch.isfinite(y).all()


@pytest.mark.parametrize("shape", [(2, 300, 420), (1, 97, 64), (3, 512, 512), (1, 1080, 1920)])
def test_preprocess_resize_normalize(shape):
    """GPU input preprocessing (dataset.py:21-24) vs the oracle on the same uint8 images: <= 2e-5 absolute."""
    from fpnmt.engine import preprocess
    n, h, w = shape
    g = torch.Generator().manual_seed(h * 7 + w)
    img = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)
    out = preprocess(img.cuda(), 512).cpu().numpy()
    for i in range(n):
        ref = O.load_image_array(img[i].numpy(), 512)
        assert np.abs(out[i] - ref).max() < 2e-5
    assert out.min() >= -1.0 - 1e-6 and out.max() <= 1.0 + 1e-6
