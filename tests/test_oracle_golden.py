"""Pin the oracle to the reference: compare it with the golden vectors produced by the REFERENCE'S OWN modules
(tests/golden/make_golden.py imports /root/reference/{models,utils,layers} on a numpy stand-in for TensorFlow and
records their outputs).  Float comparisons are fp32-vs-fp32 of the same arithmetic in a different summation order:
tolerance 2e-5 relative L2 for layer-level outputs, 2e-4 for the whole CNN + encoder stack; token ids are exact.
"""
import os

import numpy as np
import pytest
import torch

import fpnmt_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
UNITS = os.path.join(HERE, "golden", "reference_units.npz")
MODEL = os.path.join(HERE, "golden", "reference_model.npz")


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.fixture(scope="module")
def units():
    with np.load(UNITS) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def model():
    with np.load(MODEL) as z:
        return {k: z[k] for k in z.files}


def sub_weights(z, prefix):
    return {k[len(prefix):].replace("|", "/"): v for k, v in z.items() if k.startswith(prefix)}


def test_positional_encoding_and_mask(units):
    assert rel(O.raw_positional_encoding(40, 64).numpy(), units["unit_pos_40_64"]) < 1e-6
    rows = O.positional_encoding(1024, 512)[0, [0, 1, 17, 255, 1023]].numpy()
    assert np.abs(rows - units["unit_pos_1024_512_rows"]).max() < 2e-6
    assert np.array_equal(O.create_look_ahead_mask(5).numpy(), units["unit_mask_5"])


def test_scaled_dot_product_attention(units):
    q, k, v = (torch.from_numpy(units["unit_sdpa_" + n]) for n in "qkv")
    out, att = O.scaled_dot_product_attention(q, k, v, None)
    assert rel(out.numpy(), units["unit_sdpa_out"]) < 2e-6 and rel(att.numpy(), units["unit_sdpa_att"]) < 2e-6
    out_m, _ = O.scaled_dot_product_attention(torch.from_numpy(units["unit_sdpa_qm"]), k, v, O.create_look_ahead_mask(5))
    assert rel(out_m.numpy(), units["unit_sdpa_out_masked"]) < 2e-6


def test_multi_head_attention(units):
    w = O.W(sub_weights(units, "unit_mha_w|"))
    q, kv = torch.from_numpy(units["unit_mha_q"]), torch.from_numpy(units["unit_mha_kv"])
    y, _ = O.mha(w, "m", kv, kv, q, None, num_heads=4)
    assert rel(y.numpy(), units["unit_mha_out"]) < 2e-5


def test_encoder_layer_queries_are_the_layer_input(units):
    """models/transformer.py:185-190: all four cross-level attentions use the layer INPUT as query (the reference's
    `out += ...` rebinds; it does not update `baseline`)."""
    w = O.W(sub_weights(units, "unit_enc_w|"))
    views = [torch.from_numpy(units["unit_enc_view%d" % i]) for i in range(5)]
    y = O.encoder_layer(views, w, "e", num_heads=4)
    assert rel(y.numpy(), units["unit_enc_out"]) < 2e-5


def test_decoder_and_transformer_call(units):
    wd = sub_weights(units, "unit_dec_w|")
    wd = {k.replace("t/", "transformer/", 1): v for k, v in wd.items()}
    w = O.W(wd)
    enc = torch.from_numpy(units["unit_dec_enc_out"])
    toks = torch.from_numpy(units["unit_dec_tokens"]).long()
    mask = O.create_look_ahead_mask(6)
    hidden, att = O.decoder(toks, enc, w, mask, max_seq_len=12, num_layers=2, num_heads=4)
    assert rel(hidden.numpy(), units["unit_dec_hidden"]) < 2e-5
    assert rel(att["decoder_layer2_block2"].numpy(), units["unit_dec_att_l2_b2"]) < 2e-5
    logits, _ = O.transformer_logits(enc, toks, w, mask, max_seq_len=12, num_layers=2, num_heads=4)
    assert rel(logits.numpy(), units["unit_dec_logits"]) < 2e-5
    y1, _, _ = O.decoder_layer(enc[:, :4], enc, w, "transformer/decoder/dec_layers/0", O.create_look_ahead_mask(4), None, 4)
    assert rel(y1.numpy(), units["unit_dec_layer0_out"]) < 2e-5


def test_coattention(units):
    s1 = torch.ones(1, 1, 7, 7)
    h1 = torch.arange(1 * 7 * 7 * 3, dtype=torch.float32).reshape(1, 7, 7, 3).permute(0, 3, 1, 2)
    out = O.coattention_cnn(s1, h1).permute(0, 2, 3, 1)
    assert rel(out.numpy(), units["unit_coatt_demo_out"]) < 1e-6          # == hs / 49 (coattention.py:44-45 demo inputs)
    s2 = torch.from_numpy(units["unit_coatt_score"]).permute(0, 3, 1, 2)
    h2 = torch.from_numpy(units["unit_coatt_hs"]).permute(0, 3, 1, 2)
    assert rel(O.coattention_cnn(s2, h2).permute(0, 2, 3, 1).numpy(), units["unit_coatt_out"]) < 2e-6


@pytest.fixture(scope="module")
def model_run(model):
    L, V, Tm, S = (int(v) for v in model["fe_cfg"])
    w = O.test_weights("mobilenet224_1.0", vocab=V, layers=L, seed=0)
    img = O.test_images(2, S, seed=int(model["fe_image_seed"][0]))
    Wv = O.W(w)
    taps = {}
    with torch.no_grad():
        mem = O.encoder(img, Wv, "mobilenet224_1.0", num_layers=L, input_vocab_size=1024, taps=taps)
    return dict(L=L, V=V, Tm=Tm, S=S, Wv=Wv, img=img, mem=mem, taps=taps)


def test_feature_extractor_and_fpn_taps(model, model_run):
    """retinanet.py:105-141, 266-307 + mobilenet.py:64-66 taps, as executed by the reference's own functional graph."""
    t = model_run["taps"]
    for n in ("C3", "C4", "C5", "P3", "P4", "P5", "P6", "P7"):
        a = t[n][:1].numpy()
        assert list(a.shape) == list(model["fe_tap_%s_shape" % n])
        assert rel(a[:, ::4, ::4, ::8], model["fe_tap_%s_sub" % n]) < 2e-4, n
        s = model["fe_tap_%s_sum" % n]
        assert abs(a.astype(np.float64).sum() - s[0]) <= 2e-4 * s[1], n
    for i in range(5):
        assert rel(t["features"][i][:1].numpy(), model["fe_feat%d" % i]) < 2e-4, i


def test_encoder_memory(model, model_run):
    assert rel(model_run["mem"].numpy(), model["fe_memory"]) < 2e-4
    assert np.abs(model_run["mem"].numpy() - model["fe_memory"]).max() < 2e-3


@pytest.mark.parametrize("beam", [4, 8])
def test_predict_token_ids_identical(model, model_run, beam):
    """utils/pipeline.py:82-154 executed by the reference vs the oracle's faithful and cached forms."""
    r = model_run
    for i in range(2):
        gold = model["predict_beam%d_img%d" % (beam, i)]
        ids = O.predict_reference(r["img"][i], r["Wv"], r["Tm"], beam, 2, 3, "mobilenet224_1.0", num_layers=r["L"],
                                  mode="prob", enc_output=r["mem"][i:i + 1])
        assert ids.tolist() == gold.tolist()
    ids_c, len_c = O.predict_batch_cached(r["mem"], r["Wv"], r["Tm"], beam, 2, 3, num_layers=r["L"])
    for i in range(2):
        gold = model["predict_beam%d_img%d" % (beam, i)]
        assert ids_c[i, :len_c[i]].tolist() == gold.tolist()


def test_predict_stops_on_end_token(model, model_run):
    r = model_run
    end = int(model["predict_end_token"][0])
    ids = O.predict_reference(r["img"][0], r["Wv"], r["Tm"], 4, 2, end, "mobilenet224_1.0", num_layers=r["L"], mode="prob",
                              enc_output=r["mem"][0:1])
    assert ids.tolist() == model["predict_beam4_img0_with_end"].tolist()


def test_teacher_forced_logits(model, model_run):
    r = model_run
    toks = torch.from_numpy(model["predict_tf_tokens"]).long()
    logits, _ = O.transformer_logits(r["mem"][1:2], toks, r["Wv"], O.create_look_ahead_mask(toks.shape[1]), r["Tm"], r["L"])
    got = logits[0, :, ::4].numpy()
    assert np.abs(got - model["predict_tf_logits_rows"]).max() < 2e-3     # the north-star bound on per-step log-probs
    assert rel(got, model["predict_tf_logits_rows"]) < 2e-4


def test_load_image_resize_and_preprocess(units):
    """dataset.py:19-26 (minus the JPEG decode) as executed by the reference vs the oracle and the host mirror."""
    from fpnmt.dataset import preprocess_input, resize_bilinear_tf2
    img = units["load_image_input_u8"]
    ref = units["load_image_out_sub"]
    got = O.load_image_array(img, 512)
    assert got.shape == (512, 512, 3) and got.dtype == np.float32
    assert np.abs(got[::3, ::3] - ref).max() < 2e-5
    assert units["load_image_out_minmax"][0] >= -1.0 and units["load_image_out_minmax"][1] <= 1.0
    host = preprocess_input(resize_bilinear_tf2(img.astype(np.float32), 512, 512))
    assert np.abs(host[::3, ::3] - ref).max() < 2e-5
    got2 = O.load_image_array(units["load_image_input2_u8"], 512)[::7, ::5]
    assert np.abs(got2 - units["load_image_out2_sub"]).max() < 2e-5
