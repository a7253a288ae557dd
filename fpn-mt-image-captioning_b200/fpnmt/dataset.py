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


class Tokenizer:
    """The subset of keras.preprocessing.text.Tokenizer the hot path uses (pipeline.py:19,89-90,169,188)."""

    def __init__(self, word_index: Dict[str, int], index_word: Optional[Dict[int, str]] = None, config: Optional[dict] = None,
                 word_counts: Optional[dict] = None, word_docs: Optional[dict] = None, index_docs: Optional[dict] = None):
        self.word_index = dict(word_index)
        self.index_word = dict(index_word) if index_word is not None else {i: w for w, i in self.word_index.items()}
        self.config = config or {}
        self.word_counts, self.word_docs, self.index_docs = word_counts or {}, word_docs or {}, index_docs or {}

    def sequences_to_texts(self, sequences: Iterable[Iterable[int]]) -> List[str]:
        out = []
        for seq in sequences:
            words = []
            for i in seq:
                w = self.index_word.get(int(i))
                if w is not None:          # Keras skips indices it does not know (0 = padding)
                    words.append(w)
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
        """Vocabulary for synthetic runs: pad=0 (no word), <unk>=1, <start>=2, <end>=3, w4..w{V-1}."""
        wi = {"<unk>": C.UNK_ID, "<start>": C.START_ID, "<end>": C.END_ID}
        for i in range(4, vocab):
            wi["w%d" % i] = i
        iw = {i: w for w, i in wi.items()}
        iw[C.PAD_ID] = "<pad>"               # dataset.py:62,67-68 adds index 0 = '<pad>' to index_word
        return cls(wi, iw)


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
