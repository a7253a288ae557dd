"""`Pipeline.predict` beam loop restated (utils/pipeline.py:82-154).  TEST INFRASTRUCTURE ONLY.

`predict_reference` is the faithful form: one image, encoder once, the decoder recomputed
over the whole prefix every step (no KV cache), scores = running PRODUCT of softmax
probabilities in float32, `tf.math.top_k` tie-break (lower flat index first), stop as soon as
the top beam emits <end>.  `beam_step` is that loop's body on raw logits (what the CUDA
decode-tail kernel is checked against).  `predict_batch_cached` is the batched, KV-cached,
log-domain form the CUDA engine implements; tests prove it token-identical to the faithful
form wherever the float32 product has not underflowed.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .model import (TR, create_look_ahead_mask, dense, encoder, layer_norm, leaky_relu, raw_positional_encoding,
                    transformer_logits)
from .ops import W

__all__ = ["top_k_stable", "beam_step", "predict_reference", "predict_batch_cached", "teacher_forced_logprobs",
           "strip_result"]


def top_k_stable(values: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """tf.math.top_k on a 1-D array: descending, equal elements ordered by lower index first."""
    order = np.argsort(-values, kind="stable")[:k]
    return values[order], order.astype(np.int64)


def beam_step(logits: np.ndarray, beam_score: np.ndarray, mode: str = "prob"):
    """One iteration of pipeline.py:115-141 for one image.

    logits (N,V) float32 = last-position decoder output; beam_score (N,) float32.
    mode "prob": score = softmax(logits) * beam_prob   (the reference, pipeline.py:117-123)
    mode "log" : score = log_softmax(logits) + beam_logprob (same ordering while no underflow)
    Returns (parent (N,), token (N,), new_score (N,)).
    """
    n, v = logits.shape
    x = torch.from_numpy(np.ascontiguousarray(logits, dtype=np.float32))
    if mode == "prob":
        cand = torch.softmax(x, dim=-1).numpy() * beam_score.astype(np.float32)[:, None]     # :117,:122
    else:
        cand = torch.log_softmax(x, dim=-1).numpy() + beam_score.astype(np.float32)[:, None]
    vals, idx = top_k_stable(cand.reshape(-1), n)                                              # :123,:128
    parent = idx // v                                                                          # :130
    token = idx - parent * v                                                                   # :131
    return parent, token, vals.astype(np.float32)


def _beam_step_finished(logits: np.ndarray, beam_score: np.ndarray, fin_len: np.ndarray, lp: np.ndarray, t: int, end_token: int):
    """One step of the finished-beam / length-penalty extension for one image (see predict_batch_cached)."""
    n, v = logits.shape
    raw = (torch.log_softmax(torch.from_numpy(np.ascontiguousarray(logits, dtype=np.float32)), dim=-1).numpy()
           + beam_score.astype(np.float32)[:, None])
    key = np.full((n, v), -np.inf, np.float32)
    for b in range(n):
        if fin_len[b] > 0:
            raw[b, :] = -np.inf
            raw[b, end_token] = beam_score[b]                              # the frozen beam itself
            key[b, end_token] = np.float32(beam_score[b]) / lp[fin_len[b]]
        else:
            key[b] = raw[b] / lp[t + 1]
    _, idx = top_k_stable(key.reshape(-1), n)
    parent = idx // v
    token = idx - parent * v
    new_fin = np.where(fin_len[parent] > 0, fin_len[parent], np.where(token == end_token, t + 1, 0))
    return parent, token, raw.reshape(-1)[idx].astype(np.float32), new_fin


def strip_result(seq: np.ndarray, end_token: int) -> np.ndarray:
    """pipeline.py:147-154: drop <start>, and the trailing <end> if present."""
    return seq[1:-1] if seq[-1] == end_token else seq[1:]


def predict_reference(img_hwc: torch.Tensor, w: W, max_seq_len: int, beam: int, start_token: int, end_token: int,
                      backbone: str = "mobilenet224_1.0", num_layers: int = 6, num_heads: int = 8,
                      mode: str = "prob", enc_output: Optional[torch.Tensor] = None,
                      trace: Optional[dict] = None) -> np.ndarray:
    """Faithful restatement of Pipeline.predict for ONE image (pipeline.py:82-154)."""
    if enc_output is None:
        enc_output = encoder(img_hwc[None], w, backbone, num_layers, num_heads)                # :93-94
    enc_output = enc_output.repeat(beam, 1, 1)                                                 # :97
    vocab = w(TR + "/final_layer/kernel").shape[1]
    beam_output = np.full((beam, 1), start_token, dtype=np.int64)                              # :101
    beam_score = np.ones((beam,), np.float32) if mode == "prob" else np.zeros((beam,), np.float32)   # :102
    result = None
    for _ in range(max_seq_len):                                                               # :105
        t = beam_output.shape[1]
        mask = create_look_ahead_mask(t, w.dtype)                                              # :106
        logits, _ = transformer_logits(enc_output, torch.from_numpy(beam_output), w, mask, max_seq_len,
                                       num_layers, num_heads)                                  # :109-112
        last = logits[:, -1, :].to(torch.float32).numpy().reshape(beam, vocab)                 # :115,:119
        parent, token, beam_score = beam_step(last, beam_score, mode)
        beam_output = np.concatenate([beam_output[parent], token[:, None]], axis=-1)           # :134-137
        best = int(np.argmax(beam_score))                                                      # :143
        result = beam_output[best]                                                             # :144
        if trace is not None:
            trace.setdefault("score", []).append(beam_score.copy())
            trace.setdefault("parent", []).append(parent.copy())
            trace.setdefault("token", []).append(token.copy())
        if result[-1] == end_token:                                                            # :147
            return result[1:-1]
    return strip_result(result, end_token)                                                     # :151-154


# ---- batched, KV-cached, log-domain form (what the engine implements) ------------------------------
def _split_heads(x, h):
    b, t, d = x.shape
    return x.reshape(b, t, h, d // h).permute(0, 2, 1, 3)


def _dec_step_cached(w: W, tok: torch.Tensor, pos_row: torch.Tensor, caches: List[dict], cross: List[dict],
                     num_layers: int, num_heads: int):
    """One decoder position for rows (R,) given self-attn caches and per-row cross K/V."""
    emb = w(TR + "/decoder/embedding/embeddings")
    x = emb[tok] + pos_row                                                                    # (R,d)
    x = x[:, None, :]
    for l in range(num_layers):
        p = TR + "/decoder/dec_layers/%d" % l
        q = dense(x, w, p + "/mha1/wq")
        k = dense(x, w, p + "/mha1/wk")
        v = dense(x, w, p + "/mha1/wv")
        c = caches[l]
        c["k"] = k if c["k"] is None else torch.cat([c["k"], k], dim=1)
        c["v"] = v if c["v"] is None else torch.cat([c["v"], v], dim=1)
        qh, kh, vh = _split_heads(q, num_heads), _split_heads(c["k"], num_heads), _split_heads(c["v"], num_heads)
        att = torch.softmax(qh @ kh.transpose(-1, -2) / np.sqrt(qh.shape[-1]), dim=-1) @ vh
        att = att.permute(0, 2, 1, 3).reshape(x.shape)
        out1 = layer_norm(dense(att, w, p + "/mha1/dense") + x, w(p + "/layernorm1/gamma"), w(p + "/layernorm1/beta"))
        q2 = _split_heads(dense(out1, w, p + "/mha2/wq"), num_heads)
        att2 = torch.softmax(q2 @ cross[l]["k"].transpose(-1, -2) / np.sqrt(q2.shape[-1]), dim=-1) @ cross[l]["v"]
        att2 = att2.permute(0, 2, 1, 3).reshape(x.shape)
        out2 = layer_norm(dense(att2, w, p + "/mha2/dense") + out1, w(p + "/layernorm2/gamma"), w(p + "/layernorm2/beta"))
        ffn = dense(leaky_relu(dense(out2, w, p + "/ffn1"), 0.2), w, p + "/ffn2")
        x = layer_norm(ffn + out2, w(p + "/layernorm3/gamma"), w(p + "/layernorm3/beta"))
    return dense(x[:, 0, :], w, TR + "/final_layer")


def predict_batch_cached(enc_output: torch.Tensor, w: W, max_seq_len: int, beam: int, start_token: int,
                         end_token: int, num_layers: int = 6, num_heads: int = 8, early_stop: bool = True,
                         trace: Optional[dict] = None, true_beam: bool = False, finished_beams: bool = False,
                         length_penalty: float = 0.0) -> Tuple[np.ndarray, np.ndarray]:
    """Batched KV-cached beam decode with log-domain scores.

    enc_output (B,16,d).  Returns (ids (B,max_seq_len) int32 padded with 0, lengths (B,) int32), each row being
    what `predict_reference` returns for that image (start stripped, trailing <end> stripped).
    An image stops the first time its top beam emits <end> (pipeline.py:147); stopped images are frozen.
    true_beam=True is the flagged EXTENSION (not reference behaviour): only beam 0 is alive at t = 0
    (log-score 0, the others -inf), so the N beams diverge instead of staying N copies of the greedy path
    (pipeline.py:101-102 starts all N identical).
    finished_beams=True / length_penalty=alpha is the second flagged EXTENSION: a beam that has emitted <end> is frozen (it
    contributes one candidate - itself, score unchanged - instead of V), candidates are ranked by score / lp(length) with
    lp(len) = ((5 + len) / 6) ** alpha (length = generated tokens including <end>), and an image stops when its BEST beam is a
    frozen one.  finished_beams=False, alpha=0 is the reference (pipeline.py:143-148: stop the moment the top beam emits <end>;
    lower beams that emitted <end> keep decoding).
    """
    if length_penalty != 0.0 and not finished_beams:
        raise ValueError("length_penalty only orders beams of different lengths: it needs finished_beams=True")
    lp = np.array([np.power(np.float32((5.0 + i) / 6.0), np.float32(length_penalty)) for i in range(max_seq_len + 2)], np.float32)
    fin_len = np.zeros((enc_output.shape[0], beam), np.int64)
    bsz, _, d = enc_output.shape
    rows = bsz * beam
    enc_rows = enc_output.repeat_interleave(beam, dim=0)
    pos = raw_positional_encoding(max_seq_len, d).to(w.dtype)
    cross = []
    for l in range(num_layers):
        p = TR + "/decoder/dec_layers/%d" % l
        cross.append({"k": _split_heads(dense(enc_rows, w, p + "/mha2/wk"), num_heads),
                      "v": _split_heads(dense(enc_rows, w, p + "/mha2/wv"), num_heads)})
    caches = [{"k": None, "v": None} for _ in range(num_layers)]
    seqs = np.full((bsz, beam, 1), start_token, dtype=np.int64)
    score = np.zeros((bsz, beam), np.float32)
    if true_beam:
        score[:, 1:] = -np.inf
    done = np.zeros((bsz,), bool)
    out_ids = np.zeros((bsz, max_seq_len), np.int32)
    out_len = np.zeros((bsz,), np.int32)
    for t in range(max_seq_len):
        tok = torch.from_numpy(seqs[:, :, -1].reshape(rows))
        logits = _dec_step_cached(w, tok, pos[t], caches, cross, num_layers, num_heads).to(torch.float32).numpy()
        logits = logits.reshape(bsz, beam, -1)
        gparent = np.zeros((bsz, beam), np.int64)
        new_seqs = np.zeros((bsz, beam, t + 2), np.int64)
        for b in range(bsz):
            if finished_beams:
                parent, token, sc, fin_len[b] = _beam_step_finished(logits[b], score[b], fin_len[b], lp, t, end_token)
            else:
                parent, token, sc = beam_step(logits[b], score[b], "log")
            gparent[b] = b * beam + parent
            new_seqs[b] = np.concatenate([seqs[b][parent], token[:, None]], axis=-1)
            score[b] = sc
            if trace is not None:
                trace.setdefault("logits", {}).setdefault(t, {})[b] = logits[b].copy()
        seqs = new_seqs
        gp = torch.from_numpy(gparent.reshape(rows))
        for c in caches:                                  # reorder the self-attention cache by beam parent
            c["k"], c["v"] = c["k"][gp], c["v"][gp]
        for b in range(bsz):
            if done[b]:
                continue
            res = seqs[b, int(np.argmax(score[b]))] if not finished_beams else seqs[b, 0]
            if finished_beams and fin_len[b, 0] > 0:            # best beam is frozen: its caption ends before its <end>
                res = res[:fin_len[b, 0] + 1]
            if res[-1] == end_token or t == max_seq_len - 1:
                r = strip_result(res, end_token)
                out_ids[b, :len(r)] = r
                out_len[b] = len(r)
                done[b] = True
        if early_stop and done.all():
            break
    return out_ids, out_len


def teacher_forced_logprobs(enc_output: torch.Tensor, tokens: torch.Tensor, w: W, max_seq_len: int,
                            num_layers: int = 6, num_heads: int = 8) -> torch.Tensor:
    """log_softmax of Transformer.call logits (uncached, whole prefix) — (B,t,V)."""
    mask = create_look_ahead_mask(tokens.shape[1], w.dtype)
    logits, _ = transformer_logits(enc_output, tokens, w, mask, max_seq_len, num_layers, num_heads)
    return torch.log_softmax(logits.to(torch.float32), dim=-1)
