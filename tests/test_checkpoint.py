"""TF2 object-graph checkpoint reader / writer (fpnmt/checkpoint.py; /root/reference/utils/pipeline.py:38-48, train.py:95-96).

No TensorFlow here, so the format is pinned by its own invariants - CRC-32C known answers (RFC 3720 vectors), LevelDB block
checksums, footer magic - by hand-assembled bytes for the prefix-compressed key path, and by round trips writer -> reader on the
real variable tree."""
import os
import struct

import numpy as np
import pytest

from fpnmt import checkpoint as ck
from fpnmt.weights import init_weights, model_spec


def test_crc32c_known_answers():
    assert ck.crc32c(b"123456789") == 0xE3069283
    assert ck.crc32c(b"\x00" * 32) == 0x8A9136AA                     # RFC 3720 B.4
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    data = np.random.default_rng(0).integers(0, 256, 1_000_003, dtype=np.uint8).tobytes()
    ref = 0
    for i in range(0, len(data), 1000):                              # the scalar walk, chained
        ref = ck.crc32c(data[i:i + 1000], ref)
    assert ck.crc32c(data) == ref                                    # the vectorised lanes + GF(2) combine agree with it
    assert ck._unmask(ck._mask(0xE3069283)) == 0xE3069283
    assert ck._mask(0xE3069283) == (((0xE3069283 >> 15) | (0xE3069283 << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_reader_follows_prefix_compressed_blocks_assembled_by_hand(tmp_path):
    """An index whose single data block is written out byte by byte here (shared-prefix entries, one restart), independent of
    write_tensor_bundle."""
    def entry(shared, key_delta, value):
        return bytes([shared, len(key_delta), len(value)]) + key_delta + value
    def block(body, restarts):
        raw = body + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))
        return raw + b"\x00" + struct.pack("<I", ck._mask(ck.crc32c(raw + b"\x00")))
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    b = np.array([7, 8], dtype=np.int64)
    data = a.tobytes() + b.tobytes()
    ea = ck._entry_proto(1, a.shape, 0, a.nbytes, ck._mask(ck.crc32c(a.tobytes())))
    eb = ck._entry_proto(9, b.shape, a.nbytes, b.nbytes, ck._mask(ck.crc32c(b.tobytes())))
    body = entry(0, b"", b"\x08\x01") + entry(0, b"net/dense/bias", eb) + entry(10, b"kernel", ea)      # "net/dense/" shared
    blk = block(body, [0])
    idx = block(entry(0, b"net/dense/kernel", bytes([0, len(blk) - 5])), [0])
    meta = block(b"", [0])
    f = blk + meta + idx
    footer = bytes([len(blk), len(meta) - 5, len(blk) + len(meta), len(idx) - 5])
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    (tmp_path / "m.index").write_bytes(f + footer)
    (tmp_path / "m.data-00000-of-00001").write_bytes(data)
    t = ck.read_tensor_bundle(str(tmp_path / "m"))
    assert sorted(t) == ["net/dense/bias", "net/dense/kernel"]
    assert np.array_equal(t["net/dense/kernel"], a) and np.array_equal(t["net/dense/bias"], b)


def test_round_trip_of_the_whole_variable_tree_and_corruption_is_caught(tmp_path):
    w = init_weights("mobilenet224_1.0", vocab=512, seed=3, num_layers=2)
    prefix = str(tmp_path / "train" / "ckpt-4")
    ck.weights_to_checkpoint(w, prefix, save_counter=4)
    assert ck.latest_checkpoint(str(tmp_path / "train")) == prefix
    header, entries = ck.read_index(prefix + ".index")
    assert header["num_shards"] == 1 and len(entries) == len(w) + 1
    key = "transformer/decoder/dec_layers/0/mha1/wq/kernel/.ATTRIBUTES/VARIABLE_VALUE"
    assert entries[key]["shape"] == [512, 512] and entries[key]["dtype"] == 1
    got = ck.load_checkpoint(str(tmp_path / "train"), "mobilenet224_1.0", num_layers=2)
    assert list(got) == [p for p, *_ in model_spec("mobilenet224_1.0", 512, 2)]
    for k in w:
        assert np.array_equal(got[k], w[k]), k
    # a name-subset read touches only those tensors
    sub = ck.read_tensor_bundle(prefix, names=[key])
    assert list(sub) == [key]
    # corruption: one flipped byte in the data shard -> that tensor's CRC; one in the index -> the block's CRC
    dpath = prefix + ".data-00000-of-00001"
    raw = bytearray(open(dpath, "rb").read())
    raw[entries[key]["offset"] + 5] ^= 0x40
    open(dpath, "wb").write(bytes(raw))
    with pytest.raises(ck.CheckpointError, match="tensor checksum"):
        ck.read_tensor_bundle(prefix, names=[key])
    assert ck.read_tensor_bundle(prefix, names=[key], verify=False)[key].shape == (512, 512)
    ipath = prefix + ".index"
    raw = bytearray(open(ipath, "rb").read())
    raw[100] ^= 0x01
    open(ipath, "wb").write(bytes(raw))
    with pytest.raises(ck.CheckpointError, match="block checksum"):
        ck.read_index(ipath)
    open(ipath, "wb").write(bytes(raw[:-8]) + b"\x00" * 8)
    with pytest.raises(ck.CheckpointError, match="magic"):
        ck.read_index(ipath)


def test_object_graph_names_of_a_tf_checkpoint_map_to_the_variable_tree():
    """Names as tf.train.Checkpoint(transformer=..., optimizer=...) writes them: optimizer slots, save_counter and the graph blob are
    dropped; functional-model layers (`layer_with_weights-N`) resolve through the layer order; bf16 / shape errors are named."""
    S = ck._SUFFIX
    spec = model_spec("mobilenet224_1.0", 512, 1)
    w = {p: np.zeros(s, np.float32) for p, s, *_ in spec}
    fe = "transformer/encoder/feature_extractor/retinanet_model"
    order = ck._default_layer_order([(p, s) for p, s, *_ in spec], fe)
    assert order[0] == "Conv1" and "P3" in order and "regression_submodel/pyramid_regression_0" in order
    t = {}
    for p, a in w.items():
        if p.startswith(fe + "/"):
            layer, var = p[len(fe) + 1:].rsplit("/", 1)
            t["%s/layer_with_weights-%d/%s%s" % (fe, order.index(layer), var, S)] = a + order.index(layer)
        else:
            t[p + S] = a
    t["optimizer/iter" + S] = np.zeros((), np.int64)
    t["transformer/final_layer/kernel/.OPTIMIZER_SLOT/optimizer/m" + S] = np.zeros((512, 512), np.float32)
    t["save_counter" + S] = np.ones((), np.int64)
    got = ck.checkpoint_to_weights(t, expected=[(p, s) for p, s, *_ in spec])
    assert list(got) == [p for p, *_ in spec]
    assert float(got[fe + "/P3/kernel"].ravel()[0]) == order.index("P3")
    bad = dict(t)
    bad["transformer/final_layer/bias" + S] = np.zeros((511,), np.float32)
    with pytest.raises(ck.CheckpointError, match="final_layer/bias"):
        ck.checkpoint_to_weights(bad, expected=[(p, s) for p, s, *_ in spec])
    del bad["transformer/final_layer/bias" + S]
    with pytest.raises(ck.CheckpointError, match="no variable"):
        ck.checkpoint_to_weights(bad, expected=[(p, s) for p, s, *_ in spec])


@pytest.mark.gpu
def test_pipeline_restores_from_a_checkpoint_manager_directory(tmp_path):
    """Pipeline(tokenizer, checkpoint_path, T) with checkpoint_path = a CheckpointManager directory (pipeline.py:38-48): the
    engine built from `checkpoint` + `ckpt-3.index/.data` emits the ids of the engine built from the same weights as .npz."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    import fpnmt_oracle as O
    from fpnmt.dataset import Tokenizer, store_tokenizer_to_path
    from fpnmt.pipeline import Pipeline
    from fpnmt.weights import save_weights
    vocab = 304
    tok_path = str(tmp_path / "_tokenizer.json")
    store_tokenizer_to_path(Tokenizer.synthetic(vocab), tok_path)
    w = O.test_weights("mobilenet224_1.0", vocab=vocab, layers=6, seed=4)
    w = {k: np.asarray(v, np.float32) for k, v in w.items()}
    ckdir, npzdir = tmp_path / "tfck", tmp_path / "npz"
    ckdir.mkdir(); npzdir.mkdir()
    ck.weights_to_checkpoint(w, str(ckdir / "ckpt-3"), save_counter=3)
    save_weights(str(npzdir / "weights.npz"), w)
    img = O.test_images(2, 512, seed=8).numpy()
    a = Pipeline(tok_path, str(ckdir), 10, beam=4)
    ids_a, len_a = a.predict_batch(img, 10)
    del a
    b = Pipeline(tok_path, str(npzdir), 10, beam=4)
    ids_b, len_b = b.predict_batch(img, 10)
    assert (ids_a == ids_b).all() and (len_a == len_b).all()
