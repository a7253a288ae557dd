"""Python owner of one `fpnmt_handle` (one per GPU).  PyTorch is used for device memory and streams only.

Tensors cross the C ABI as raw device pointers: torch CUDA tensors directly, anything else that speaks
DLPack (`__dlpack__`) through `torch.from_dlpack`.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import config as cfg


def _as_cuda(x, dtype, device) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        if hasattr(x, "__dlpack__"):
            x = torch.from_dlpack(x)
        else:
            x = torch.as_tensor(np.asarray(x))
    if x.device.type != "cuda":
        x = x.to(device, non_blocking=True)
    if x.dtype != dtype:
        x = x.to(dtype)
    return x.contiguous()


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Engine:
    """B200 caption-inference engine for a fixed (backbone, batch, beam, vocab, max_len)."""

    def __init__(self, weights: Dict[str, np.ndarray], backbone: str = "mobilenet224_1.0", batch: int = 1,
                 beam: int = cfg.BEAM_SEARCH_N, vocab: Optional[int] = None, max_len: int = cfg.SYNTH_MAX_SEQ_LEN,
                 num_layers: int = cfg.num_layers, d_model: int = cfg.d_model, num_heads: int = cfg.num_heads,
                 dff: int = cfg.dff, image_size: int = cfg.IMAGE_INPUT_SIZE, precision: str = "bf16",
                 score_mode: str = "log", start_id: int = cfg.START_ID, end_id: int = cfg.END_ID,
                 true_beam: bool = False, use_graphs: bool = True, device: int = 0, opts: Sequence[str] = (),
                 cache_mode: str = "ancestry", decode_path: str = "auto", length_penalty: float = 0.0,
                 finished_beams: bool = False, dec_groups: int = 0, lanes: int = 1, _exp: int = 0):
        """opts: names from _lib.OPT_BITS (e.g. "no_xattn") - each turns one fused kernel back into its unfused equivalent;
        cache_mode "ancestry" | "physical"; decode_path "auto" (today: the per-operator chain) | "chain" | "fused" (one
        dstep_kernel launch runs every layer, the vocabulary projection and the beam tail of all steps); length_penalty / finished_beams: flagged extensions, 0 = reference;
        lanes: batches in flight for `generate_stream` (each lane is a complete engine on the same GPU; 2 overlaps the encoder
        of batch i+1 with the decode of batch i)."""
        if not torch.cuda.is_available():
            raise RuntimeError("fpnmt.Engine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        if vocab is None:
            vocab = int(weights["transformer/final_layer/kernel"].shape[1])
        self.device = torch.device("cuda", device)
        self.backbone, self.batch, self.beam, self.vocab, self.max_len = backbone, batch, beam, vocab, max_len
        self.image_size, self.d_model, self.num_layers = image_size, d_model, num_layers
        self.precision = precision
        c = _lib.FpnmtConfig()
        c.backbone = _lib.BACKBONE_IDS[backbone]
        c.image_size, c.batch, c.beam, c.vocab, c.max_len = image_size, batch, beam, vocab, max_len
        c.num_layers, c.d_model, c.num_heads, c.dff = num_layers, d_model, num_heads, dff
        c.precision = _lib.PREC_IDS[precision]
        c.score_mode = _lib.SCORE_IDS[score_mode]
        c.start_id, c.end_id = start_id, end_id
        c.true_beam, c.use_graphs = int(true_beam), int(use_graphs)
        c.kernel_opts = 0
        for o in opts:
            c.kernel_opts |= _lib.OPT_BITS[o]
        c.cache_mode, c.decode_path = _lib.CACHE_IDS[cache_mode], _lib.DECODE_IDS[decode_path]
        c.length_penalty, c.finished_beams, c.dec_groups = float(length_penalty), int(finished_beams), int(dec_groups)
        c.lanes = int(lanes)
        self.lanes = max(1, int(lanes))
        c.reserved[0] = int(_exp)          # developer A/B switches of the fused decoder (dstep.cuh), 0 in product use
        self._h = C.c_void_p()
        torch.cuda.init()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fpnmt_create(C.byref(c), device, C.byref(self._h)))
            for key, arr in weights.items():
                a = np.ascontiguousarray(arr, dtype=np.float32)
                shape = (C.c_int64 * a.ndim)(*a.shape)
                _lib.check(self.lib.fpnmt_set_weight(self._h, key.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim))
            _lib.check(self.lib.fpnmt_finalize_weights(self._h))
        side = image_size // 16
        self.feature_shapes = [(batch, side >> i, side >> i, d_model) for i in range(5)]
        self.n_memory = (side >> 3) ** 2

    # ------------------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self.lib.fpnmt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _images(self, images) -> Tuple[int, int, object]:
        """Returns (pointer, on_host, keepalive)."""
        s = self.image_size
        if isinstance(images, torch.Tensor) and images.device.type == "cpu" or isinstance(images, np.ndarray):
            t = torch.as_tensor(images)
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(torch.float32).contiguous()
            if tuple(t.shape) != (self.batch, s, s, 3):
                raise ValueError("images must be NHWC (%d,%d,%d,3), got %s" % (self.batch, s, s, tuple(t.shape)))
            return t.data_ptr(), 1, t
        t = _as_cuda(images, torch.float32, self.device)
        if tuple(t.shape) != (self.batch, s, s, 3):
            raise ValueError("images must be NHWC (%d,%d,%d,3), got %s" % (self.batch, s, s, tuple(t.shape)))
        return t.data_ptr(), 0, t

    def encode(self, images) -> torch.Tensor:
        """Encoder.call (transformer.py:266-303): (B,S,S,3) -> (B, n_memory, d_model) float32 on device."""
        ptr, on_host, keep = self._images(images)
        out = torch.empty((self.batch, self.n_memory, self.d_model), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.fpnmt_encode(self._h, ptr, on_host, out.data_ptr(), _stream_ptr(self.device)))
        return out

    def features(self, images) -> List[torch.Tensor]:
        """FeatureExtractor.call (retinanet.py:306-307): five NHWC float32 maps."""
        ptr, on_host, keep = self._images(images)
        outs = [torch.empty(sh, dtype=torch.float32, device=self.device) for sh in self.feature_shapes]
        arr = (C.c_void_p * 5)(*[o.data_ptr() for o in outs])
        _lib.check(self.lib.fpnmt_features(self._h, ptr, on_host, arr, _stream_ptr(self.device)))
        return outs

    def tap(self, name: str) -> torch.Tensor:
        n = C.c_size_t(0)
        _lib.check(self.lib.fpnmt_get_tap(self._h, name.encode(), None, 0, C.byref(n), _stream_ptr(self.device)))
        out = torch.empty((n.value,), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.fpnmt_get_tap(self._h, name.encode(), out.data_ptr(), n.value, C.byref(n),
                                          _stream_ptr(self.device)))
        return out

    def decode_logits(self, memory, tokens) -> torch.Tensor:
        """Transformer.call(enc_output, tar, False, mask) (transformer.py:359-374) -> logits (B,t,V)."""
        tok = _as_cuda(tokens, torch.int32, self.device)
        t = int(tok.shape[1])
        mem_ptr = None
        if memory is not None:
            mem = _as_cuda(memory, torch.float32, self.device)
            mem_ptr = mem.data_ptr()
        out = torch.empty((self.batch, t, self.vocab), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.fpnmt_decode_logits(self._h, mem_ptr, tok.data_ptr(), t, out.data_ptr(),
                                                _stream_ptr(self.device)))
        return out

    def decode_hidden(self, memory, tokens) -> torch.Tensor:
        """Decoder.call(x, enc_output, False, mask, None) (transformer.py:321-341) -> last decoder layer's output (B,t,d)."""
        tok = _as_cuda(tokens, torch.int32, self.device)
        t = int(tok.shape[1])
        mem_ptr = None
        if memory is not None:
            mem = _as_cuda(memory, torch.float32, self.device)
            mem_ptr = mem.data_ptr()
        out = torch.empty((self.batch, t, self.d_model), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.fpnmt_decode_hidden(self._h, mem_ptr, tok.data_ptr(), t, out.data_ptr(), _stream_ptr(self.device)))
        return out

    def beam_step(self, logits, scores) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """One decode-tail step (pipeline.py:115-141) on caller logits (B*N,V) and scores (B*N,)."""
        lg = _as_cuda(logits, torch.float32, self.device)
        sc = _as_cuda(scores, torch.float32, self.device)
        rows = self.batch * self.beam
        parent = torch.empty((rows,), dtype=torch.int32, device=self.device)
        token = torch.empty((rows,), dtype=torch.int32, device=self.device)
        out = torch.empty((rows,), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.fpnmt_beam_step(self._h, lg.data_ptr(), sc.data_ptr(), parent.data_ptr(), token.data_ptr(),
                                            out.data_ptr(), _stream_ptr(self.device)))
        return parent, token, out

    def generate(self, images, early_stop: bool = True, to_host: bool = True, return_scores: bool = False):
        """Pipeline.predict for a batch (pipeline.py:82-154): ids (B,T) int32 zero-padded, lengths (B,)."""
        ptr, on_host, keep = self._images(images)
        if to_host:
            ids = torch.empty((self.batch, self.max_len), dtype=torch.int32).pin_memory()
            lens = torch.empty((self.batch,), dtype=torch.int32).pin_memory()
        else:
            ids = torch.empty((self.batch, self.max_len), dtype=torch.int32, device=self.device)
            lens = torch.empty((self.batch,), dtype=torch.int32, device=self.device)
        scores = None
        if return_scores:
            scores = torch.zeros((self.max_len, self.batch), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.fpnmt_generate(self._h, ptr, on_host, ids.data_ptr(), lens.data_ptr(), int(to_host),
                                           int(early_stop), scores.data_ptr() if scores is not None else None,
                                           _stream_ptr(self.device)))
        return (ids, lens, scores) if return_scores else (ids, lens)

    def stage(self, host_images, slot: int):
        """Enqueue the host->device copy of one batch into staging slot 0/1 on the engine's copy stream (returns at once).
        Returns the host tensor actually handed to the copy; keep it alive until the matching generate_staged is done."""
        t = torch.as_tensor(host_images)
        if t.device.type != "cpu":
            raise ValueError("stage() takes HOST images; device images go straight to generate()")
        s = self.image_size
        if tuple(t.shape) != (self.batch, s, s, 3):
            raise ValueError("images must be NHWC (%d,%d,%d,3), got %s" % (self.batch, s, s, tuple(t.shape)))
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(torch.float32).contiguous()
        _lib.check(self.lib.fpnmt_stage_images(self._h, t.data_ptr(), int(slot)))
        return t

    def generate_staged(self, slot: int, early_stop: bool = True, to_host: bool = True):
        """`generate` on a batch previously handed to `stage(…, slot)`."""
        if to_host:
            ids = torch.empty((self.batch, self.max_len), dtype=torch.int32).pin_memory()
            lens = torch.empty((self.batch,), dtype=torch.int32).pin_memory()
        else:
            ids = torch.empty((self.batch, self.max_len), dtype=torch.int32, device=self.device)
            lens = torch.empty((self.batch,), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.fpnmt_generate_staged(self._h, int(slot), ids.data_ptr(), lens.data_ptr(), int(to_host),
                                                  int(early_stop), None, _stream_ptr(self.device)))
        return ids, lens

    def submit(self, lane: int, images, early_stop: bool = False):
        """Enqueue one whole batch on lane `lane` (fpnmt_submit) and return at once.  Returns a keep-alive object: the images
        must stay valid until `collect(lane)`."""
        ptr, on_host, keep = self._images_nocopy(images)
        _lib.check(self.lib.fpnmt_submit(self._h, int(lane), ptr, on_host, int(early_stop), _stream_ptr(self.device)))
        return keep

    def collect(self, lane: int, to_host: bool = True):
        """Result of the batch submitted on `lane` (fpnmt_collect): ids (B,T) int32 zero-padded, lengths (B,)."""
        if to_host:
            ids = torch.empty((self.batch, self.max_len), dtype=torch.int32).pin_memory()
            lens = torch.empty((self.batch,), dtype=torch.int32).pin_memory()
        else:
            ids = torch.empty((self.batch, self.max_len), dtype=torch.int32, device=self.device)
            lens = torch.empty((self.batch,), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.fpnmt_collect(self._h, int(lane), ids.data_ptr(), lens.data_ptr(), int(to_host),
                                          _stream_ptr(self.device)))
        return ids, lens

    def _images_nocopy(self, images):
        """(pointer, on_host, keepalive) for submit: host batches go as they are (the lane copies them on its own stream),
        device batches are used in place."""
        return self._images(images)

    def generate_stream(self, batches, early_stop: bool = True, to_host: bool = True):
        """Captions for an iterable of batches (HOST tensors: pinned memory makes the copies asynchronous; or device tensors),
        yielded as (ids, lens) per batch, in order.  With `lanes` >= 2 the batches go round-robin through the lanes
        (fpnmt_submit / fpnmt_collect): host->device copy, encoder and decode of up to `lanes` batches are in flight at once, so
        the encoder of batch i+1 runs under the decode of batch i; results are identical to `generate` batch by batch.  With one
        lane: the double-buffered input copy of fpnmt_stage_images / fpnmt_generate_staged (the tf.data prefetch of
        dataset.py:92)."""
        it = iter(batches)
        if self.lanes >= 2:
            L, pending, i = self.lanes, [], 0
            for b in it:
                if len(pending) == L:
                    lane, _keep = pending.pop(0)
                    yield self.collect(lane, to_host=to_host)
                lane = i % L
                pending.append((lane, self.submit(lane, b, early_stop=early_stop)))
                i += 1
            for lane, _keep in pending:
                yield self.collect(lane, to_host=to_host)
            return
        try:
            first = next(it)
        except StopIteration:
            return
        if isinstance(first, torch.Tensor) and first.device.type == "cuda":
            b = first
            while b is not None:
                yield self.generate(b, early_stop=early_stop, to_host=to_host)
                b = next(it, None)
            return
        keep = [self.stage(first, 0), None]
        i = 0
        while True:
            nxt = next(it, None)
            if nxt is not None:
                keep[(i + 1) & 1] = self.stage(nxt, (i + 1) & 1)
            yield self.generate_staged(i & 1, early_stop=early_stop, to_host=to_host)
            if nxt is None:
                return
            i += 1

    def decode(self, early_stop: bool = True, to_host: bool = True):
        """Decode half only, from the memory left by the last `encode`."""
        if to_host:
            ids = torch.empty((self.batch, self.max_len), dtype=torch.int32).pin_memory()
            lens = torch.empty((self.batch,), dtype=torch.int32).pin_memory()
        else:
            ids = torch.empty((self.batch, self.max_len), dtype=torch.int32, device=self.device)
            lens = torch.empty((self.batch,), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.fpnmt_decode(self._h, ids.data_ptr(), lens.data_ptr(), int(to_host), int(early_stop), None,
                                         _stream_ptr(self.device)))
        return ids, lens

    def profile(self, iters: int = 5) -> dict:
        buf = C.create_string_buffer(1 << 20)
        _lib.check(self.lib.fpnmt_profile(self._h, iters, buf, len(buf)))
        return json.loads(buf.value.decode())

    @property
    def launch_count(self) -> int:
        return int(self.lib.fpnmt_launch_count(self._h))


def conv2d(x: torch.Tensor, kernel_hwio: np.ndarray, bias: Optional[np.ndarray] = None, act: int = 0,
           pad: Tuple[int, int] = (0, 0), residual: Optional[torch.Tensor] = None, res_mode: int = 0,
           precision: str = "bf16", force_bn: int = 0) -> torch.Tensor:
    """Stand-alone implicit-GEMM convolution (stride 1) through the C ABI — used by the kernel parity tests."""
    lib = _lib.load()
    dev = x.device
    x = x.to(torch.float32).contiguous()
    n, h, w, cin = x.shape
    k = np.ascontiguousarray(kernel_hwio, dtype=np.float32)
    kh, kw, _, cout = k.shape
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    out = torch.empty((n, h, w, cout), dtype=torch.float32, device=dev)
    r = None if residual is None else residual.to(torch.float32).contiguous()
    _lib.check(lib.fpnmt_op_conv2d(dev.index or 0, _lib.PREC_IDS[precision], x.data_ptr(), n, h, w, cin,
                                   k.ctypes.data_as(C.c_void_p), kh, kw, cout, pad[0], pad[1],
                                   None if b is None else b.ctypes.data_as(C.c_void_p), act,
                                   None if r is None else r.data_ptr(), res_mode, out.data_ptr(), force_bn,
                                   _stream_ptr(dev)))
    return out


def dense(x: torch.Tensor, kernel: np.ndarray, bias: Optional[np.ndarray] = None, act: int = 0,
          residual: Optional[torch.Tensor] = None, gamma: Optional[np.ndarray] = None, beta: Optional[np.ndarray] = None,
          eps: float = 1e-6, precision: str = "bf16", force_bn: int = 0) -> torch.Tensor:
    """Stand-alone skinny-row Dense (+ residual + LayerNorm) through the C ABI — used by the kernel parity tests."""
    lib = _lib.load()
    dev = x.device
    x = x.to(torch.float32).contiguous()
    r, k = x.shape
    kn = np.ascontiguousarray(kernel, dtype=np.float32)
    f = kn.shape[1]
    host = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    b, g, be = host(bias), host(gamma), host(beta)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    res = None if residual is None else residual.to(torch.float32).contiguous()
    out = torch.empty((r, f), dtype=torch.float32, device=dev)
    _lib.check(lib.fpnmt_op_dense(dev.index or 0, _lib.PREC_IDS[precision], x.data_ptr(), r, k, ptr(kn), f, ptr(b), act,
                                  None if res is None else res.data_ptr(), ptr(g), ptr(be), eps, out.data_ptr(), force_bn,
                                  _stream_ptr(dev)))
    return out


def preprocess(images_u8: torch.Tensor, size: int = cfg.IMAGE_INPUT_SIZE) -> torch.Tensor:
    """GPU `load_image` after the JPEG decode (dataset.py:21-24): uint8 (N,H,W,3) -> float32 (N,size,size,3) in [-1,1]."""
    lib = _lib.load()
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
        raise ValueError("images must be uint8 (N,H,W,3)")
    if images_u8.device.type != "cuda":
        raise RuntimeError("fpnmt.preprocess needs CUDA tensors; there is no CPU fallback")
    x = images_u8.contiguous()
    n, h, w, _ = x.shape
    out = torch.empty((n, size, size, 3), dtype=torch.float32, device=x.device)
    _lib.check(lib.fpnmt_op_preprocess(x.device.index or 0, x.data_ptr(), n, h, w, size, out.data_ptr(), _stream_ptr(x.device)))
    return out


def decode_jpeg(jpegs: Sequence[bytes], size: int = cfg.IMAGE_INPUT_SIZE, device: int = 0, return_sizes: bool = False):
    """GPU `load_image` for a batch of encoded JPEG files (dataset.py:19-26): nvJPEG decode (channels=3) -> bilinear resize
    (TF2 half-pixel, no antialias) -> x / 127.5 - 1.  Returns float32 (n, size, size, 3) on the device, ready for
    `Engine.encode` / `generate`; the decoded RGB images never visit the host."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("fpnmt.decode_jpeg needs a CUDA device; there is no CPU fallback")
    n = len(jpegs)
    bufs = [np.frombuffer(b, dtype=np.uint8) for b in jpegs]
    ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
    lens = (C.c_size_t * n)(*[b.size for b in bufs])
    dev = torch.device("cuda", device)
    out = torch.empty((n, size, size, 3), dtype=torch.float32, device=dev)
    sizes = np.zeros((n, 2), np.int32)
    with torch.cuda.device(dev):
        _lib.check(lib.fpnmt_op_decode_jpeg(device, ptrs, lens, n, size, out.data_ptr(), sizes.ctypes.data_as(C.c_void_p),
                                            _stream_ptr(dev)))
        torch.cuda.current_stream(dev).synchronize()       # the encoded bytes (host) may be released after this
    return (out, sizes) if return_sizes else out
