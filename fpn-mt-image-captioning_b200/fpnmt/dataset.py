"""Host-side data formats either side of the hot path (mirror of /root/reference/dataset.py).

* `load_image`         dataset.py:19-26   file -> (512,512,3) float32 in [-1,1]  (PIL decode + TF2-style bilinear)
* `Tokenizer` + `load_tokenizer_from_path` / `store_tokenizer_to_path`   dataset.py:96-146 (Keras Tokenizer json,
  double-encoded: `json.dumps(tokenizer.to_json())`)
* `store_additional_info` / `load_additional_info`   dataset.py:248-258
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Optional

import numpy as np

from . import config as C


def resize_bilinear_tf2(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """tf.image.resize(img, (h,w)) defaults: bilinear, half-pixel centres, antialias=False.  img (H,W,C) float."""
    h, w, _ = img.shape
    ys = (np.arange(out_h, dtype=np.float64) + 0.5) * (h / out_h) - 0.5
    xs = (np.arange(out_w, dtype=np.float64) + 0.5) * (w / out_w) - 0.5
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    wy = (ys - y0).astype(np.float32)[:, None, None]
    wx = (xs - x0).astype(np.float32)[None, :, None]
    y0c, y1c = np.clip(y0, 0, h - 1), np.clip(y0 + 1, 0, h - 1)
    x0c, x1c = np.clip(x0, 0, w - 1), np.clip(x0 + 1, 0, w - 1)
    img = img.astype(np.float32)
    top = img[y0c][:, x0c] * (1 - wx) + img[y0c][:, x1c] * wx
    bot = img[y1c][:, x0c] * (1 - wx) + img[y1c][:, x1c] * wx
    return top * (1 - wy) + bot * wy


def preprocess_input(img: np.ndarray) -> np.ndarray:
    """tf.keras.applications.mobilenet_v2.preprocess_input: x / 127.5 - 1."""
    return img.astype(np.float32) / 127.5 - 1.0


def load_image(img_path: str, caption=None):
    """dataset.py:19-26 — returns (img (S,S,3) float32 in [-1,1], caption)."""
    from PIL import Image
    with Image.open(img_path) as im:
        arr = np.asarray(im.convert("RGB"), dtype=np.float32)
    arr = resize_bilinear_tf2(arr, C.IMAGE_INPUT_SIZE, C.IMAGE_INPUT_SIZE)
    return preprocess_input(arr), caption


KERAS_DEFAULT_FILTERS = '!"#$%&()*+,-./:;<=>?@[\\]^_`{|}~\t\n'
REFERENCE_FILTERS = '!"#$%&()*+-/:;=?@[\\]^_`{|}~ '      # dataset.py:60 (keeps '.', ',', '<', '>')


def load_images_gpu(img_paths, size: int = C.IMAGE_INPUT_SIZE, device: int = 0):
    """dataset.py:19-26 for a list of JPEG files on the GPU: read_file on the host, decode_jpeg + resize + preprocess_input on
    the device (nvJPEG + k_preprocess).  Returns float32 (n, size, size, 3) on the device."""
    from .engine import decode_jpeg
    data = []
    for p in img_paths:
        with open(p, "rb") as f:
            data.append(f.read())
    return decode_jpeg(data, size=size, device=device)


class Tokenizer:
    """The part of keras.preprocessing.text.Tokenizer the hot path uses (pipeline.py:19,89-90,169,188), with Keras'
    semantics for the two constructor arguments the reference sets (dataset.py:58-60, restored from the JSON config by
    dataset.py:113): an id >= `num_words`, and an id that is not in `index_word` while `oov_token` is set, both print the
    OOV word; without an OOV token such ids are dropped.  Pinned by tests/golden/tokenizer_golden.json, a file written by
    the reference's own `store_tokenizer_to_path`."""

    def __init__(self, word_index: Dict[str, int], index_word: Optional[Dict[int, str]] = None, config: Optional[dict] = None,
                 word_counts: Optional[dict] = None, word_docs: Optional[dict] = None, index_docs: Optional[dict] = None):
        self.word_index = dict(word_index)
        self.index_word = dict(index_word) if index_word is not None else {i: w for w, i in self.word_index.items()}
        self.config = dict(config or {})
        self.num_words = self.config.get("num_words")
        self.oov_token = self.config.get("oov_token")
        self.filters = self.config.get("filters", KERAS_DEFAULT_FILTERS)
        self.lower = self.config.get("lower", True)
        self.split = self.config.get("split", " ")
        self.word_counts, self.word_docs, self.index_docs = word_counts or {}, word_docs or {}, index_docs or {}

    def _words(self, text: str) -> List[str]:
        if self.lower:
            text = text.lower()
        text = text.translate(str.maketrans({c: self.split for c in self.filters}))
        return [w for w in text.split(self.split) if w]

    def texts_to_sequences(self, texts: Iterable[str]) -> List[List[int]]:
        oov = self.word_index.get(self.oov_token)
        out = []
        for text in texts:
            vect = []
            for w in self._words(text):
                i = self.word_index.get(w)
                if i is not None:
                    if self.num_words and i >= self.num_words:
                        if oov is not None:
                            vect.append(oov)
                    else:
                        vect.append(i)
                elif self.oov_token is not None:
                    vect.append(oov)
            out.append(vect)
        return out

    def sequences_to_texts(self, sequences: Iterable[Iterable[int]]) -> List[str]:
        oov = self.word_index.get(self.oov_token)
        out = []
        for seq in sequences:
            words = []
            for i in seq:
                i = int(i)
                w = self.index_word.get(i)
                if w is not None:
                    if self.num_words and i >= self.num_words:
                        if oov is not None:
                            words.append(self.index_word[oov])
                    else:
                        words.append(w)
                elif self.oov_token is not None:
                    words.append(self.index_word[oov])
            out.append(" ".join(words))
        return out

    def to_json(self) -> str:
        cfg = dict(self.config)
        cfg["word_counts"] = json.dumps(self.word_counts)
        cfg["word_docs"] = json.dumps(self.word_docs)
        cfg["index_docs"] = json.dumps({str(k): v for k, v in self.index_docs.items()})
        cfg["word_index"] = json.dumps(self.word_index)
        cfg["index_word"] = json.dumps({str(k): v for k, v in self.index_word.items()})
        return json.dumps({"class_name": "Tokenizer", "config": cfg})

    @classmethod
    def synthetic(cls, vocab: int) -> "Tokenizer":
        """Vocabulary for synthetic runs, shaped like the reference's (dataset.py:58-65): '' = 0 (padding), the OOV word
        "unk" = 1, <start> = 2, <end> = 3, w4..w{V-1}; num_words = TOP_K, oov_token = "unk"."""
        wi = {"unk": C.UNK_ID, "<start>": C.START_ID, "<end>": C.END_ID}
        for i in range(4, vocab):
            wi["w%d" % i] = i
        iw = {i: w for w, i in wi.items()}
        wi[""] = C.PAD_ID                    # dataset.py:64-65
        iw[C.PAD_ID] = ""
        cfg = {"num_words": C.TOP_K, "filters": REFERENCE_FILTERS, "lower": True, "split": " ", "char_level": False,
               "oov_token": "unk", "document_count": 0}
        return cls(wi, iw, cfg)


def _tokenizer_from_json(json_string: str) -> Tokenizer:
    """dataset.py:96-123 (keras `tokenizer_from_json`)."""
    tokenizer_config = json.loads(json_string)
    config = tokenizer_config.get("config")
    word_counts = json.loads(config.pop("word_counts", "{}"))
    word_docs = json.loads(config.pop("word_docs", "{}"))
    index_docs = {int(k): v for k, v in json.loads(config.pop("index_docs", "{}")).items()}
    index_word = {int(k): v for k, v in json.loads(config.pop("index_word")).items()}
    word_index = json.loads(config.pop("word_index"))
    return Tokenizer(word_index, index_word, config, word_counts, word_docs, index_docs)


def load_tokenizer_from_path(path: str) -> Tokenizer:
    """dataset.py:125-135 — the file holds json.dumps(tokenizer.to_json()), i.e. a JSON string of a JSON document."""
    with open(path) as f:
        data = json.load(f)
    if isinstance(data, dict):            # tolerate a singly-encoded file
        data = json.dumps(data)
    return _tokenizer_from_json(data)


def store_tokenizer_to_path(tokenizer: Tokenizer, path: str) -> None:
    """dataset.py:137-146."""
    with open(path, "w", encoding="utf-8") as f:
        f.write(json.dumps(tokenizer.to_json(), ensure_ascii=False))


def store_additional_info(info: dict, path: str) -> None:
    with open(path, "w") as f:
        json.dump(info, f)


def load_additional_info(path: str) -> dict:
    try:                                   # dataset.py:252-258: a missing/corrupt file yields {}
        with open(path) as f:
            return json.load(f)
    except Exception:
        return {}
