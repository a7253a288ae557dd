#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE'S OWN Python modules.  TEST INFRASTRUCTURE ONLY.

    python tests/golden/make_golden.py            # needs /root/reference (build container only)

TensorFlow cannot be installed here, so /root/reference/{models,utils,layers,common,...} are imported on top of
`tests/golden/tfshim/tensorflow` (a numpy stand-in for the TF primitives they call — see its docstring) and executed
unmodified.  What is recorded, each with the inputs and the weights that produced it:

  unit_*        get_angles / raw_positional_encoding / create_look_ahead_mask / scaled_dot_product_attention /
                MultiHeadAttention / EncoderLayer / DecoderLayer / Decoder / Transformer.call(training=False) /
                CoAttention_CNN  (models/transformer.py, models/coattention.py) at a small width (d=64)
  fe_*          FeatureExtractor.call and Encoder.call (models/retinanet.py:266-307, models/transformer.py:266-303) on a
                256x256 image: the reference's own FPN wiring, head truncation and MobileNetV2 taps
  predict_*     Pipeline.predict (utils/pipeline.py:82-154), beam 4 (the reference constant) and beam 8

Weights come from `fpnmt_oracle.test_weights` (seeded `fpnmt.weights.init_weights` + fixed gains + calibrated BatchNorm
statistics; images from `fpnmt.synthetic.structured_images`) and are assigned to the reference objects by walking the
attribute / layer-name paths of SURVEY.md Appendix B, so the same dict drives the oracle and the CUDA engine.
The fixtures are committed; the GPU box never runs this script.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FPNMT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "fpn-mt-image-captioning_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "tfshim"))

import tensorflow as tf  # noqa: E402  (the shim)

assert tf.__version__.endswith("numpy-shim")

import fpnmt_oracle as O  # noqa: E402   (only for the shared test-weight recipe: gains + BatchNorm calibration)

tf.reset_uids()
import models.transformer as T  # noqa: E402   (reference module, unmodified)
import models.coattention as CO  # noqa: E402
import utils.pipeline as P  # noqa: E402


def assign_tree(obj, prefix: str, w: dict) -> int:
    """Assign every weight whose key starts with `prefix` to the reference object graph rooted at `obj`."""
    n = 0
    for key, val in w.items():
        if not key.startswith(prefix + "/"):
            continue
        parts = key[len(prefix) + 1:].split("/")
        cur = obj
        for p in parts[:-1]:
            if isinstance(cur, (list, tuple)):
                cur = cur[int(p)]
            elif isinstance(cur, tf.keras.Model) and cur.functional:
                cands = [l for l in cur.all_layers() if l.name == p]
                assert len(cands) == 1, (key, p, [l.name for l in cur.all_layers()][:8])
                cur = cands[0]
            else:
                cur = getattr(cur, p)
        cur.assign(parts[-1], val)
        n += 1
    return n


def small_weights(rng, names_shapes):
    return {k: rng.normal(0, s, shp).astype(np.float32) for k, shp, s in names_shapes}


def mha_spec(name, d):
    out = []
    for p in ("wq", "wk", "wv", "dense"):
        out += [(name + "/%s/kernel" % p, (d, d), d ** -0.5), (name + "/%s/bias" % p, (d,), 0.1)]
    return out


def ln_spec(name, d):
    return [(name + "/gamma", (d,), 0.2), (name + "/beta", (d,), 0.1)]


def unit_goldens(out: dict):
    rng = np.random.default_rng(7)
    d, h, dff, V, L, T_max = 64, 4, 128, 50, 2, 12
    # positional encoding / mask
    out["unit_pos_40_64"] = np.asarray(T.raw_positional_encoding(40, d))
    out["unit_pos_1024_512_rows"] = np.asarray(T.positional_encoding(1024, 512))[0, [0, 1, 17, 255, 1023]]
    out["unit_mask_5"] = np.asarray(T.create_look_ahead_mask(5))
    # scaled dot product attention
    q = rng.normal(size=(2, h, 3, 16)).astype(np.float32)
    k = rng.normal(size=(2, h, 5, 16)).astype(np.float32)
    v = rng.normal(size=(2, h, 5, 16)).astype(np.float32)
    o, a = T.scaled_dot_product_attention(tf.constant(q), tf.constant(k), tf.constant(v), None)
    out.update(unit_sdpa_q=q, unit_sdpa_k=k, unit_sdpa_v=v, unit_sdpa_out=np.asarray(o), unit_sdpa_att=np.asarray(a))
    qm = rng.normal(size=(2, h, 5, 16)).astype(np.float32)
    om, _ = T.scaled_dot_product_attention(tf.constant(qm), tf.constant(k), tf.constant(v), T.create_look_ahead_mask(5))
    out.update(unit_sdpa_qm=qm, unit_sdpa_out_masked=np.asarray(om))
    # MultiHeadAttention
    w = small_weights(rng, mha_spec("m", d))
    m = T.MultiHeadAttention(d, h)
    assign_tree(m, "m", w)
    xq = rng.normal(size=(2, 3, d)).astype(np.float32)
    xk = rng.normal(size=(2, 7, d)).astype(np.float32)
    y, _ = m(tf.constant(xk), tf.constant(xk), tf.constant(xq), None)
    out.update({"unit_mha_w|" + kk.replace("/", "|"): vv for kk, vv in w.items()})
    out.update(unit_mha_q=xq, unit_mha_kv=xk, unit_mha_out=np.asarray(y))
    # EncoderLayer (5 views, baseline last)
    spec = []
    for i in range(4):
        spec += mha_spec("e/mhas/%d" % i, d)
    spec += [("e/ffn1/kernel", (d, dff), d ** -0.5), ("e/ffn1/bias", (dff,), 0.1), ("e/ffn2/kernel", (dff, d), dff ** -0.5),
             ("e/ffn2/bias", (d,), 0.1)] + ln_spec("e/layernorm1", d) + ln_spec("e/layernorm2", d)
    w = small_weights(rng, spec)
    for kk in w:
        if kk.endswith("gamma"):
            w[kk] = (w[kk] + 1).astype(np.float32)
    el = T.EncoderLayer(d, h, dff)
    assign_tree(el, "e", w)
    views = [rng.normal(size=(2, n, d)).astype(np.float32) for n in (9, 6, 4, 2, 3)]
    y = el([tf.constant(x) for x in views], False, None)
    out.update({"unit_enc_w|" + kk.replace("/", "|"): vv for kk, vv in w.items()})
    out.update({"unit_enc_view%d" % i: x for i, x in enumerate(views)})
    out["unit_enc_out"] = np.asarray(y)
    # Transformer.call(training=False): Decoder (embedding + pos + L DecoderLayers) + final_layer
    spec = [("t/decoder/embedding/embeddings", (V, d), 0.5)]
    for l in range(L):
        p = "t/decoder/dec_layers/%d" % l
        spec += mha_spec(p + "/mha1", d) + mha_spec(p + "/mha2", d)
        spec += [(p + "/ffn1/kernel", (d, dff), d ** -0.5), (p + "/ffn1/bias", (dff,), 0.1),
                 (p + "/ffn2/kernel", (dff, d), dff ** -0.5), (p + "/ffn2/bias", (d,), 0.1)]
        spec += ln_spec(p + "/layernorm1", d) + ln_spec(p + "/layernorm2", d) + ln_spec(p + "/layernorm3", d)
    spec += [("t/final_layer/kernel", (d, V), d ** -0.5), ("t/final_layer/bias", (V,), 0.1)]
    w = small_weights(rng, spec)
    for kk in w:
        if kk.endswith("gamma"):
            w[kk] = (w[kk] + 1).astype(np.float32)
    dec = T.Decoder(L, d, h, dff, V, 0.1, 0, T_max)
    final = tf.keras.layers.Dense(V, activation="linear")
    holder = type("H", (), {})()
    holder.decoder, holder.final_layer = dec, final
    assign_tree(holder, "t", w)
    enc_out = rng.normal(size=(3, 5, d)).astype(np.float32)
    toks = rng.integers(0, V, size=(3, 6)).astype(np.int32)
    mask = T.create_look_ahead_mask(6)
    x, att = dec(tf.constant(toks), tf.constant(enc_out), False, mask, None)
    logits = final(x)
    y1, _, _ = dec.dec_layers[0](tf.constant(enc_out[:, :4]), tf.constant(enc_out), False, T.create_look_ahead_mask(4), None)
    out.update({"unit_dec_w|" + kk.replace("/", "|"): vv for kk, vv in w.items()})
    out.update(unit_dec_enc_out=enc_out, unit_dec_tokens=toks, unit_dec_hidden=np.asarray(x), unit_dec_logits=np.asarray(logits),
               unit_dec_layer0_out=np.asarray(y1),
               unit_dec_att_l2_b2=np.asarray(att["decoder_layer2_block2"]))
    # CoAttention_CNN: the reference's own demo inputs (coattention.py:44-45) and a random case
    co = CO.CoAttention_CNN()
    s1 = np.ones((1, 7, 7, 1), np.float32)
    h1 = np.arange(1 * 7 * 7 * 3, dtype=np.float32).reshape(1, 7, 7, 3)
    out["unit_coatt_demo_out"] = np.asarray(co(tf.constant(s1), tf.constant(h1)))
    s2 = rng.normal(size=(2, 4, 6, 1)).astype(np.float32) * 3
    h2 = rng.normal(size=(2, 4, 6, 5)).astype(np.float32)
    out.update(unit_coatt_score=s2, unit_coatt_hs=h2, unit_coatt_out=np.asarray(co(tf.constant(s2), tf.constant(h2))))


def build_transformer(num_layers, vocab, max_seq_len, w):
    tf.reset_uids()
    tr = T.Transformer(num_layers, 512, 8, 2048, 1024, vocab, 0.1, max_seq_len=max_seq_len)
    n = assign_tree(tr, "transformer", w)
    assert n == len(w), "assigned %d of %d weights" % (n, len(w))
    return tr


def model_goldens(out: dict):
    L, V, Tm, S = 2, 512, 10, 256
    w = O.test_weights("mobilenet224_1.0", vocab=V, layers=L, seed=0)
    tr = build_transformer(L, V, Tm, w)
    img = O.test_images(2, S, seed=1).numpy()
    fe = tr.encoder.feature_extractor
    feats = fe(tf.constant(img[:1]))
    for i, f in enumerate(feats):
        out["fe_feat%d" % i] = np.asarray(f)
    # taps of the reference's own graph: C3/C4/C5 and P3..P7 (sub-sampled to keep the fixture small)
    rn = fe.retinanet_model
    tap_model = tf.keras.Model(rn.inputs, [rn.get_layer(n).output for n in
                                           ("block_5_add", "block_12_add", "out_relu", "P3", "P4", "P5", "P6", "P7")])
    taps = tap_model(tf.constant(img[:1]))
    for n, t in zip(("C3", "C4", "C5", "P3", "P4", "P5", "P6", "P7"), taps):
        a = np.asarray(t)
        out["fe_tap_%s_shape" % n] = np.array(a.shape)
        out["fe_tap_%s_sub" % n] = a[:, ::4, ::4, ::8].copy()
        out["fe_tap_%s_sum" % n] = np.array([a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()])
    mem = np.concatenate([np.asarray(tr.encoder(tf.constant(img[i:i + 1]), False, None)) for i in range(2)], 0)
    out["fe_memory"] = mem
    out["fe_image_seed"] = np.array([1])
    out["fe_cfg"] = np.array([L, V, Tm, S])
    # Pipeline.predict, without running Pipeline.__init__ (it needs tokenizer/COCO files on disk)
    pipe = object.__new__(P.Pipeline)
    pipe.transformer = tr
    pipe.target_vocab_size = V
    pipe.max_seq_len = Tm
    tok = type("Tok", (), {})()
    tok.word_index = {"<start>": 2, "<end>": 3}
    pipe.tokenizer = tok
    for beam in (4, 8):
        P.BEAM_SEARCH_N = beam            # module constant star-imported from common_definitions.py:22
        for i in range(2):
            ids, _att = pipe.predict(tf.constant(img[i]), Tm)
            out["predict_beam%d_img%d" % (beam, i)] = np.asarray(ids).astype(np.int32)
    # the same loop with an <end> that actually fires: declare the 3rd generated token of image 0 to be <end>
    P.BEAM_SEARCH_N = 4
    ref_ids = out["predict_beam4_img0"]
    tok.word_index = {"<start>": 2, "<end>": int(ref_ids[2])}
    ids, _ = pipe.predict(tf.constant(img[0]), Tm)
    out["predict_end_token"] = np.array([int(ref_ids[2])])
    out["predict_beam4_img0_with_end"] = np.asarray(ids).astype(np.int32)
    # teacher-forced logits of Transformer.call on the generated prefix (what the per-step log-prob parity uses)
    toks = np.concatenate([[2], out["predict_beam8_img1"]]).astype(np.int32)[None, :Tm]
    logits, _ = tr(tf.constant(mem[1:2]), tf.constant(toks), False, T.create_look_ahead_mask(toks.shape[1]))
    out["predict_tf_tokens"] = toks
    out["predict_tf_logits_rows"] = np.asarray(logits)[0, :, ::4].copy()


def load_image_golden(out: dict):
    """dataset.load_image (dataset.py:19-26) as written, with the two file-system / JPEG primitives replaced by an
    in-memory uint8 array (the shim has no JPEG decoder); what is pinned is resize -> preprocess_input and dtypes."""
    import dataset as D   # reference module
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, size=(300, 420, 3), dtype=np.uint8)
    tf.io.read_file = lambda path: path
    tf.image.decode_jpeg = lambda data, channels=3: tf.constant(img)
    res, cap = D.load_image("synthetic.jpg", "a caption")
    assert cap == "a caption"
    out["load_image_input_u8"] = img
    out["load_image_out_sub"] = np.asarray(res).astype(np.float32)[::3, ::3].copy()
    out["load_image_out_minmax"] = np.array([float(np.asarray(res).min()), float(np.asarray(res).max())], np.float32)
    # an upsampling case as well (smaller than the target)
    img2 = rng.integers(0, 256, size=(97, 64, 3), dtype=np.uint8)
    tf.image.decode_jpeg = lambda data, channels=3: tf.constant(img2)
    res2, _ = D.load_image("synthetic2.jpg", None)
    out["load_image_input2_u8"] = img2
    out["load_image_out2_sub"] = np.asarray(res2).astype(np.float32)[::7, ::5].copy()


def tokenizer_golden():
    """Tokenizer wire format and detokenisation, written and read back by the reference's OWN dataset.py functions
    (`store_tokenizer_to_path` :137-146, `load_tokenizer_from_path` / `_tokenizer_from_json` :96-135) around the Keras
    Tokenizer (shim restatement).  The tokenizer is built exactly as dataset.py:58-68 does (num_words, oov_token="unk", the
    reference's filter string, '' added as index 0), with a SMALL num_words so that ids >= num_words exist.  Outputs:
    tests/golden/tokenizer_golden.json (the file the reference writes) and tokenizer_expect.json (sequences -> texts)."""
    import json
    import re
    import dataset as D   # reference module
    caps = ["the heart is normal in size .", "no acute cardiopulmonary abnormality .", "the lungs are clear , no effusion .",
            "heart size is normal . the lungs are clear .", "no pneumothorax or pleural effusion is seen .",
            "stable cardiomegaly , mild edema .", "the mediastinum is normal .", "no focal consolidation ."]
    captions = ["<start> " + c + " <end>" for c in caps]                                              # dataset.py:52
    tokenizer = tf.keras.preprocessing.text.Tokenizer(num_words=24, oov_token="unk",
                                                      filters='!"#$%&()*+-/:;=?@[\\]^_`{|}~ ')           # dataset.py:58-60
    tokenizer.fit_on_texts(captions)                                                                  # :61
    tokenizer.word_index[''] = 0                                                                      # :64
    tokenizer.index_word[0] = ''                                                                      # :65
    path = os.path.join(HERE, "tokenizer_golden.json")
    D.store_tokenizer_to_path(tokenizer, path)                                                        # :67
    tok2 = D.load_tokenizer_from_path(path)
    captions2 = [re.sub(r'([.,])', r" \1 ", c) for c in captions]                                     # :70
    seqs = tok2.texts_to_sequences(captions2)                                                         # :73
    probes = seqs[:4] + [[2, 5, 23, 24, 25, 30, 3], [0, 0, 7, 999, 1], [len(tok2.index_word) - 1, 4, 0], []]
    expect = {"num_words": tok2.num_words, "oov_token": tok2.oov_token, "vocab_size": len(tok2.index_word),
              "start_id": tok2.word_index["<start>"], "end_id": tok2.word_index["<end>"],
              "sequences": probes, "texts": tok2.sequences_to_texts(probes),
              "captions": captions2[:4], "caption_sequences": seqs[:4]}
    with open(os.path.join(HERE, "tokenizer_expect.json"), "w") as f:
        json.dump(expect, f, indent=1)
    print("tokenizer golden: vocab", expect["vocab_size"], "start", expect["start_id"], "end", expect["end_id"])


def main():
    unit, model = {}, {}
    tokenizer_golden()
    unit_goldens(unit)
    load_image_golden(unit)
    np.savez_compressed(os.path.join(HERE, "reference_units.npz"), **unit)
    model_goldens(model)
    np.savez_compressed(os.path.join(HERE, "reference_model.npz"), **model)
    for n in ("reference_units.npz", "reference_model.npz"):
        print(n, os.path.getsize(os.path.join(HERE, n)) // 1024, "KiB")


if __name__ == "__main__":
    main()
