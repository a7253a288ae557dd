"""CPU tests of the host side: the C ABI library loads and exports every symbol include/fpnmt.h declares,
the variable tree, the data formats either side of the path, multi-process sharding (gloo, world_size 2),
and that the product fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "fpnmt.h")).read() + open(os.path.join(ROOT, "include", "fpnmt_dlpack.h")).read()
    declared = sorted(set(re.findall(r"FPNMT_API\s+[\w\s\*]+?\b(fpnmt_\w+)\s*\(", hdr)))
    assert len(declared) >= 16
    lib = C.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), "libfpnmt.so does not export %s" % name
    from fpnmt import _lib
    assert sorted(_lib.SIGNATURES) == declared, "ctypes prototypes and header out of sync"


C_PROBE = r"""
/* A host with no Python and no torch: dlopen the library, bind by name, walk the error paths that need no GPU. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include "fpnmt_dlpack.h"
int main(int argc, char** argv) {
  void* so = dlopen(argv[1], RTLD_NOW);
  if (!so) { printf("dlopen: %s\n", dlerror()); return 2; }
  const char* names[] = {"fpnmt_version", "fpnmt_last_error", "fpnmt_create", "fpnmt_destroy", "fpnmt_set_weight",
    "fpnmt_finalize_weights", "fpnmt_encode", "fpnmt_features", "fpnmt_decode_logits", "fpnmt_decode_hidden", "fpnmt_beam_step",
    "fpnmt_generate", "fpnmt_submit", "fpnmt_collect", "fpnmt_lanes", "fpnmt_set_weight_dl", "fpnmt_encode_dl",
    "fpnmt_generate_dl", "fpnmt_allgather_ids", "fpnmt_allgather_ids_dl", "fpnmt_comm_create", "fpnmt_comm_unique_id",
    "fpnmt_op_decode_jpeg", "fpnmt_op_preprocess"};
  for (unsigned i = 0; i < sizeof names / sizeof *names; ++i)
    if (!dlsym(so, names[i])) { printf("missing %s\n", names[i]); return 3; }
  const char* (*version)(void) = (const char* (*)(void))dlsym(so, "fpnmt_version");
  const char* (*last_error)(void) = (const char* (*)(void))dlsym(so, "fpnmt_last_error");
  int (*create)(const fpnmt_config*, int, fpnmt_handle**) = (int (*)(const fpnmt_config*, int, fpnmt_handle**))dlsym(so, "fpnmt_create");
  int (*generate_dl)(fpnmt_handle*, const DLTensor*, DLTensor*, DLTensor*, int, DLTensor*, void*) =
      (int (*)(fpnmt_handle*, const DLTensor*, DLTensor*, DLTensor*, int, DLTensor*, void*))dlsym(so, "fpnmt_generate_dl");
  int (*allgather)(fpnmt_comm*, const int32_t*, const int32_t*, int, int, int32_t*, int32_t*, void*) =
      (int (*)(fpnmt_comm*, const int32_t*, const int32_t*, int, int, int32_t*, int32_t*, void*))dlsym(so, "fpnmt_allgather_ids");
  if (!strstr(version(), "sm_100a")) return 4;
  fpnmt_handle* h = NULL;
  if (create(NULL, 0, &h) != FPNMT_ERR_INVALID) return 5;
  fpnmt_config cfg; memset(&cfg, 0, sizeof cfg);
  int rc = create(&cfg, 0, &h);                 /* no GPU: ERR_CUDA "no CPU fallback"; GPU: ERR_INVALID (all-zero config) */
  if (rc != FPNMT_ERR_CUDA && rc != FPNMT_ERR_INVALID) return 6;
  if (!strlen(last_error())) return 7;
  if (generate_dl(NULL, NULL, NULL, NULL, 0, NULL, NULL) != FPNMT_ERR_INVALID) return 8;
  if (allgather(NULL, NULL, NULL, 1, 1, NULL, NULL, NULL) != FPNMT_ERR_INVALID) return 9;
  printf("ok %d %zu\n", rc, sizeof(fpnmt_config));
  return 0;
}
"""


def test_c_host_binds_the_library_without_python(built_lib, tmp_path):
    """A plain C program (gcc, dlopen) sees the ABI include/*.h declares: symbols, struct size, error codes and messages."""
    src = tmp_path / "probe.c"
    src.write_text(C_PROBE)
    exe = tmp_path / "probe"
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-ldl"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), built_lib], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    from fpnmt import _lib
    assert r.stdout.split()[0] == "ok" and int(r.stdout.split()[2]) == C.sizeof(_lib.FpnmtConfig)   # ctypes mirror == C struct


def test_library_has_blackwell_code(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "sm_100" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):   # tcgen05.mma, TMA load, TMA store, tcgen05.ld
        assert mnemonic in sass, "expected %s in the SASS of libfpnmt.so" % mnemonic
    # legacy mma.sync (HMMA) is allowed only where a dimension of the product is 16 (the 16 baseline queries / 16 memory
    # tokens), below tcgen05's minimum M of 64: the encoder's flash attention and the once-per-batch cross-attention operand
    # folding; every other GEMM-shaped op must be on tcgen05
    fn, offenders = None, set()
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        elif "HMMA." in line and "UTCHMMA" not in line:
            offenders.add(fn)
    # ... and the decode self-attention, one query per (row, head) against that row's own cache (M = 1): issue-bound on CUDA cores
    assert all("k_enc_attention_mma" in f or "k_xattn_fold" in f or "k_dec_self_attention_mma" in f for f in offenders), offenders


def test_no_gpu_fails_loudly(built_lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fpnmt import _lib
    lib = _lib.load()
    cfg, h = _lib.FpnmtConfig(), C.c_void_p()
    rc = lib.fpnmt_create(C.byref(cfg), 0, C.byref(h))
    assert rc == _lib.ERR_CUDA and b"no CPU fallback" in lib.fpnmt_last_error()
    from fpnmt.engine import Engine
    with pytest.raises(RuntimeError):
        Engine({}, batch=1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fpn-mt-image-captioning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "fpnmt_oracle" not in src and "oracle/" not in src, f


def test_config_mirrors_reference_constants():
    from fpnmt import config as cfg
    assert (cfg.num_layers, cfg.d_model, cfg.dff, cfg.num_heads) == (6, 512, 2048, 8)
    assert (cfg.IMAGE_INPUT_SIZE, cfg.BEAM_SEARCH_N, cfg.TOP_K) == (512, 4, 10000)
    assert (cfg.NUM_OF_PYRAMIDS, cfg.BASELINE_INDEX, cfg.N_CONV_SUBMODULE, cfg.NUM_OF_RETINANET_FILTERS) == (5, 3, 2, 256)


@pytest.mark.parametrize("backbone,nparams", [("mobilenet224_1.0", None), ("resnet50", None), ("densenet121", None)])
def test_variable_tree(backbone, nparams):
    from fpnmt.weights import model_spec, BACKBONE_TAPS
    spec = model_spec(backbone, vocab=10000)
    keys = [k for k, _, _ in spec]
    assert len(keys) == len(set(keys))
    d = dict((k, s) for k, s, _ in spec)
    assert d["transformer/encoder/enc_layers/5/mhas/3/wq/kernel"] == (512, 512)
    assert d["transformer/decoder/dec_layers/0/mha2/dense/bias"] == (512,)
    assert d["transformer/decoder/embedding/embeddings"] == (10000, 512)
    assert d["transformer/final_layer/kernel"] == (512, 10000)
    assert d["transformer/encoder/feature_extractor/model/conv2d_2/kernel"] == (3, 3, 256, 1)
    assert d["transformer/encoder/feature_extractor/model/conv2d_5/kernel"] == (3, 3, 256, 512)
    c5 = BACKBONE_TAPS[backbone][2]
    assert d["transformer/encoder/feature_extractor/retinanet_model/C5_reduced/kernel"] == (1, 1, c5, 256)


def test_init_weights_distributions_and_roundtrip(tmp_path):
    from fpnmt.weights import init_weights, load_weights, save_weights
    w = init_weights("mobilenet224_1.0", vocab=64, num_layers=1, seed=3)
    k = w["transformer/encoder/enc_layers/0/ffn1/kernel"]                      # he_normal, fan_in 512
    assert abs(k.std() - np.sqrt(2 / 512)) < 0.003 and np.abs(k).max() <= 2 * np.sqrt(2 / 512) / 0.8796 + 1e-6
    r = w["transformer/encoder/feature_extractor/retinanet_model/regression_submodel/pyramid_regression_0/kernel"]
    assert abs(r.std() - 0.01) < 5e-4
    e = w["transformer/decoder/embedding/embeddings"]
    assert e.min() >= -0.05 and e.max() <= 0.05
    assert (w["transformer/final_layer/bias"] == 0).all()
    p = str(tmp_path / "w.npz")
    save_weights(p, w)
    w2 = load_weights(p)
    assert set(w2) == set(w) and all((w2[k] == w[k]).all() for k in w)
    w3 = init_weights("mobilenet224_1.0", vocab=64, num_layers=1, seed=3)
    assert all((w3[k] == w[k]).all() for k in w)                                # deterministic


def test_tokenizer_roundtrip_double_encoded_json(tmp_path):
    from fpnmt.dataset import Tokenizer, load_tokenizer_from_path, store_tokenizer_to_path
    tok = Tokenizer.synthetic(40)
    p = str(tmp_path / "_tokenizer.json")
    store_tokenizer_to_path(tok, p)
    assert isinstance(json.load(open(p)), str)                 # json.dumps(tokenizer.to_json()) — dataset.py:143-146
    t2 = load_tokenizer_from_path(p)
    assert t2.word_index == tok.word_index and t2.index_word == tok.index_word
    assert t2.word_index["<start>"] == 2 and t2.word_index["<end>"] == 3 and len(t2.index_word) == 40
    assert t2.num_words == 10000 and t2.oov_token == "unk"     # restored from the JSON config (dataset.py:113)
    assert t2.sequences_to_texts([[5, 6, 0, 7]]) == ["w5 w6  w7"]      # padding id 0 is the empty word (dataset.py:64-65)
    assert t2.sequences_to_texts([[5, 999]]) == ["w5 unk"]     # Keras: unknown ids print the OOV word when oov_token is set
    no_oov = Tokenizer({"a": 1, "b": 2})
    assert no_oov.sequences_to_texts([[1, 7, 2]]) == ["a b"]   # ... and are dropped when it is not


def test_tokenizer_matches_file_and_texts_written_by_the_reference():
    """tests/golden/tokenizer_golden.json was written by the reference's own store_tokenizer_to_path (dataset.py:137-146)
    from a Keras Tokenizer built as dataset.py:58-65 does (make_golden.py::tokenizer_golden); tokenizer_expect.json holds
    what that tokenizer answers.  ids >= num_words and ids missing from index_word must print the OOV word."""
    from fpnmt.dataset import load_tokenizer_from_path
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    tok = load_tokenizer_from_path(os.path.join(g, "tokenizer_golden.json"))
    exp = json.load(open(os.path.join(g, "tokenizer_expect.json")))
    assert tok.num_words == exp["num_words"] and tok.oov_token == exp["oov_token"]
    assert len(tok.index_word) == exp["vocab_size"]                                # pipeline.py:19 target_vocab_size
    assert tok.word_index["<start>"] == exp["start_id"] and tok.word_index["<end>"] == exp["end_id"]   # pipeline.py:89-90
    assert tok.sequences_to_texts(exp["sequences"]) == exp["texts"]
    assert tok.texts_to_sequences(exp["captions"]) == exp["caption_sequences"]
    assert any("unk" in t for t in exp["texts"])                                   # the fixture does exercise the OOV paths


def test_load_image_contract(tmp_path):
    from PIL import Image
    from fpnmt.dataset import load_image, resize_bilinear_tf2
    rng = np.random.default_rng(0)
    arr = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    p = str(tmp_path / "a.png")
    Image.fromarray(arr).save(p)
    img, cap = load_image(p, "x")
    assert img.shape == (512, 512, 3) and img.dtype == np.float32 and cap == "x"
    assert img.min() >= -1.0 and img.max() <= 1.0
    # identity resize and exact 2x checks of the half-pixel bilinear kernel
    a = rng.random((4, 4, 1)).astype(np.float32)
    assert np.allclose(resize_bilinear_tf2(a, 4, 4), a)
    up = resize_bilinear_tf2(a, 8, 8)
    assert np.allclose(up[0, 0], a[0, 0]) and np.allclose(up[1, 1, 0], (9 * a[0, 0, 0] + 3 * a[0, 1, 0] + 3 * a[1, 0, 0] + a[1, 1, 0]) / 16)
    ref = torch.nn.functional.interpolate(torch.from_numpy(a).permute(2, 0, 1)[None], size=(7, 5), mode="bilinear",
                                          align_corners=False, antialias=False)[0].permute(1, 2, 0).numpy()
    assert np.allclose(resize_bilinear_tf2(a, 7, 5), ref, atol=1e-6)


def test_builder_signatures_match_reference():
    import inspect
    from fpnmt import retinanet as R
    from fpnmt.pipeline import Pipeline
    from fpnmt.transformer import Transformer
    assert list(inspect.signature(Pipeline.__init__).parameters)[:4] == ["self", "tokenizer_filename", "checkpoint_path", "max_seq_len"]
    assert list(inspect.signature(Pipeline.predict).parameters) == ["self", "img", "max_seq_len", "plot_layer"]
    assert list(inspect.signature(Transformer.__init__).parameters)[:10] == [
        "self", "num_layers", "d_model", "num_heads", "dff", "input_vocab_size", "target_vocab_size", "rate", "max_position", "max_seq_len"]
    assert list(inspect.signature(R.retinanet).parameters)[:7] == [
        "inputs", "backbone_layers", "num_classes", "num_anchors", "create_pyramid_features", "submodels", "name"]
    assert list(inspect.signature(R.mobilenet_retinanet).parameters)[:4] == ["num_classes", "backbone", "inputs", "modifier"]
    assert R.mobilenet_retinanet(80).backbone_layers == ("block_5_add", "block_12_add", "out_relu")
    assert R.backbone("resnet50").retinanet(80).backbone == "resnet50"
    assert R.densenet_retinanet(80).backbone_layers[2] == "conv5_block16_concat"
    with pytest.raises(ValueError):
        R.resnet_retinanet(80, backbone="resnet18")
    with pytest.raises(NotImplementedError):
        R.backbone("vgg16")


def test_shard_bounds_cover_and_balance():
    from fpnmt.dist import shard_bounds
    for total in (0, 1, 7, 64, 100):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "fpn-mt-image-captioning_b200"))
import torch
from fpnmt import dist as fd
rank, local, world = fd.init_from_env("gloo")
b, t = 3, 5
ids = (torch.arange(b * t, dtype=torch.int32).reshape(b, t) + 100 * rank)
lens = torch.full((b,), rank + 1, dtype=torch.int32)
allids, alllens = fd.allgather_captions(ids, lens, world)
assert allids.shape == (world * b, t) and alllens.shape == (world * b,)
for r in range(world):
    assert (allids[r * b:(r + 1) * b] == torch.arange(b * t, dtype=torch.int32).reshape(b, t) + 100 * r).all()
    assert (alllens[r * b:(r + 1) * b] == r + 1).all()
assert fd.max_over_ranks(float(rank)) == world - 1
lo, hi = fd.shard_bounds(10, rank, world)
sys.stdout.write("rank %%d ok %%d %%d\n" %% (rank, lo, hi))   # one write per rank: lines cannot interleave
sys.stdout.flush()
'''


def test_allgather_world_size_2_gloo(tmp_path):
    import socket
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT})
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 ok 0 5" in r.stdout and "rank 1 ok 5 10" in r.stdout


def test_bench_reference_arm_line_shape():
    import bench
    assert bench.WORKLOADS["c2"]["backbone"] == "resnet50" and bench.WORKLOADS["c2"]["batch"] == 64
    p = bench.measured_peaks()
    assert p["hbm_gbs"] > 1000 and p["tf"] > 100


def test_python_mirror_of_the_config_struct_and_option_bits_matches_the_header():
    """fpnmt/_lib.py restates `fpnmt_config` (ctypes) and the FPNMT_OPT_* bits by hand; a drift would silently select other kernels
    or shift every field.  Field names / order and every option bit are compared with include/fpnmt.h."""
    import re
    from fpnmt import _lib
    hdr = open(os.path.join(ROOT, "include", "fpnmt.h")).read()
    body = hdr[hdr.index("typedef struct fpnmt_config {"):hdr.index("} fpnmt_config;")]
    fields = re.findall(r"^\s*(?:int32_t|float)\s+(\w+)(?:\[\d+\])?;", body, flags=re.M)
    assert fields == [n for n, _ in _lib.FpnmtConfig._fields_], (fields, [n for n, _ in _lib.FpnmtConfig._fields_])
    bits = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"FPNMT_OPT_(\w+)\s*=\s*(\d+)", hdr)}
    assert bits == _lib.OPT_BITS, (sorted(bits.items()), sorted(_lib.OPT_BITS.items()))
    assert len(set(bits.values())) == len(bits) and all(v & (v - 1) == 0 for v in bits.values())      # distinct single bits
