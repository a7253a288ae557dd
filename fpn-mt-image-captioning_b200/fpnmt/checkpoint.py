"""TensorFlow-free reader (and writer) of TF2 object-graph checkpoints -> the engine's weight dict.

The reference restores `tf.train.Checkpoint(transformer=..., optimizer=...)` through a `CheckpointManager`
(/root/reference/utils/pipeline.py:38-48) and writes one every few epochs (/root/reference/train.py:95-96).  On disk that is

    <dir>/checkpoint                       text proto: model_checkpoint_path: "ckpt-7"      (CheckpointState)
    <dir>/ckpt-7.index                     TensorBundle index: a LevelDB-format SSTable, key = tensor name,
                                           value = BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}
                                           (key "" holds BundleHeaderProto {num_shards, endianness, version})
    <dir>/ckpt-7.data-00000-of-00001       the tensors' raw little-endian bytes at [offset, offset + size)

with object-graph tensor names `<attribute path>/.ATTRIBUTES/VARIABLE_VALUE`, e.g.
`transformer/decoder/dec_layers/0/mha1/wq/kernel/.ATTRIBUTES/VARIABLE_VALUE` (Python attribute names and list indices, SURVEY.md
Appendix B), optimizer slots under `.../.OPTIMIZER_SLOT/optimizer/{m,v}/...` and the serialized object graph itself under
`_CHECKPOINTABLE_OBJECT_GRAPH`.

Formats restated from their public definitions (LevelDB `table_format.md`; tensorflow/core/protobuf/tensor_bundle.proto,
tensorflow/core/util/tensor_bundle; protobuf wire format) - no TensorFlow, protobuf schema or leveldb code is used.  TensorFlow
cannot be installed in this image, so no checkpoint written by TensorFlow itself was available: the reader is pinned by the
formats' own invariants (block CRC32C of every SSTable block, per-tensor CRC32C, footer magic) and by round trips through the
writer below, which emits the same layout (prefix-compressed keys, restart points, 4 KiB data blocks, index + metaindex blocks,
48-byte footer).  What cannot be checked here is stated where it applies: the `layer_with_weights-N` numbering of Keras
functional sub-models (see `checkpoint_to_weights`).  Keras `.h5` files (models/retinanet.py:277-278) are NOT read: that needs
an HDF5 parser and there is neither h5py nor a sample file to hold one to.
"""
from __future__ import annotations

import os
import re
import struct
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57                    # LevelDB table magic
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_STRING, _DT_BFLOAT16 = 7, 14
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


class CheckpointError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ crc32c (Castagnoli)
def _make_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return np.array(tab, dtype=np.uint32)


_CRC_TABLE = _make_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC-32C of `data`.  Byte-at-a-time table walk in Python for short inputs; for tensor payloads the 8-way sliced numpy
    version below keeps a 100 MB checkpoint to a few seconds."""
    if len(data) > 4096:
        return _crc32c_np(np.frombuffer(data, dtype=np.uint8), crc)
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _gf2_times(mat: List[int], vec: int) -> int:
    s, i = 0, 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_square(mat: List[int]) -> List[int]:
    return [_gf2_times(mat, mat[n]) for n in range(32)]


_SHIFT_CACHE: Dict[int, List[int]] = {}


def _crc_shift_matrix(len2: int) -> List[int]:
    """GF(2) matrix that advances a CRC register over `len2` zero bytes (zlib's crc32_combine, Castagnoli polynomial): column n
    is the image of bit n.  Cached per length - the lanes of one buffer and most tensors of a checkpoint share theirs."""
    m = _SHIFT_CACHE.get(len2)
    if m is not None:
        return m
    odd = [0x82F63B78] + [1 << n for n in range(31)]      # one zero bit
    even = _gf2_square(odd)                               # two
    odd = _gf2_square(even)                               # four
    cols = [1 << n for n in range(32)]
    n = len2
    while True:
        even = _gf2_square(odd)
        if n & 1:
            cols = [_gf2_times(even, c) for c in cols]
        n >>= 1
        if not n:
            break
        odd = _gf2_square(even)
        if n & 1:
            cols = [_gf2_times(odd, c) for c in cols]
        n >>= 1
        if not n:
            break
    if len(_SHIFT_CACHE) < 4096:
        _SHIFT_CACHE[len2] = cols
    return cols


def _crc_combine(crc1: int, crc2: int, len2: int) -> int:
    """crc(A + B) from crc(A), crc(B), len(B)."""
    if len2 <= 0:
        return crc1
    return _gf2_times(_crc_shift_matrix(len2), crc1) ^ crc2


def _crc32c_np(a: np.ndarray, crc: int = 0) -> int:
    """Vectorised: the buffer is cut into 256 equal lanes whose CRCs advance together (one table gather per byte position),
    then the lane CRCs are merged with the GF(2) combine."""
    n = a.size
    lanes = 256
    per = n // lanes
    out = crc
    if per >= 64:
        body = a[:per * lanes].reshape(lanes, per)
        c = np.full(lanes, 0xFFFFFFFF, dtype=np.uint32)
        tab = _CRC_TABLE
        for j in range(per):
            c = tab[(c ^ body[:, j]) & 0xFF] ^ (c >> np.uint32(8))
        c ^= np.uint32(0xFFFFFFFF)
        first = True
        for v in c.tolist():
            if first:      # lane 0 continues from `crc`: crc(init=crc) == combine(crc, crc_from_zero)
                out = _crc_combine(out, int(v), per) if out else int(v)
                first = False
            else:
                out = _crc_combine(out, int(v), per)
        rest = a[per * lanes:]
    else:
        rest = a
    c = out ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in rest.tolist():
        c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _mask(crc: int) -> int:          # leveldb / TF "masked" crc stored on disk
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _unmask(m: int) -> int:
    r = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varints / protobuf wire
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


def _put_varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf: bytes) -> Iterable[Tuple[int, int, object]]:
    """(field number, wire type, value) of one protobuf message: varint -> int, fixed32/64 -> int, length-delimited -> bytes."""
    pos = 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]
            if len(v) != n:
                raise CheckpointError("truncated protobuf field")
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError("unsupported protobuf wire type %d" % wt)
        yield field, wt, v


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= 1 << 63 else v


def _parse_entry(buf: bytes) -> dict:
    """BundleEntryProto: 1 dtype, 2 shape {2: dim {1: size}}, 3 shard_id, 4 offset, 5 size, 6 crc32c (fixed32), 7 slices."""
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for f, _, v in _proto_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            for f2, _, v2 in _proto_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = _signed(v3)
                    e["shape"].append(size)
                elif f2 == 3 and v2:
                    raise CheckpointError("tensor of unknown rank in the checkpoint")
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["sliced"] = True
    return e


def _entry_proto(dtype: int, shape: Sequence[int], offset: int, size: int, crc: int) -> bytes:
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(s) for s in shape))
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


# ------------------------------------------------------------------------------------------------ SSTable
def _read_block(f: bytes, offset: int, size: int, verify: bool) -> bytes:
    raw = f[offset:offset + size + 5]
    if len(raw) != size + 5:
        raise CheckpointError("SSTable block runs past the end of the index file")
    body, ctype, stored = raw[:size], raw[size], struct.unpack_from("<I", raw, size + 1)[0]
    if verify and _unmask(stored) != crc32c(raw[:size + 1]):
        raise CheckpointError("SSTable block checksum mismatch at offset %d" % offset)
    if ctype == 1:
        raise CheckpointError("snappy-compressed SSTable block (TensorFlow writes checkpoint indexes uncompressed)")
    if ctype != 0:
        raise CheckpointError("unknown SSTable block type %d" % ctype)
    return body


def _block_entries(block: bytes) -> Iterable[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise CheckpointError("SSTable block too small")
    nrestarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestarts
    if end < 0:
        raise CheckpointError("bad restart array")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key):
            raise CheckpointError("bad key prefix length")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(index_path: str, verify: bool = True) -> Tuple[dict, Dict[str, dict]]:
    """(header, {tensor name: entry}) of a TensorBundle `.index` file."""
    with open(index_path, "rb") as fh:
        f = fh.read()
    if len(f) < 48:
        raise CheckpointError("%s: too small for an SSTable footer" % index_path)
    footer = f[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
        raise CheckpointError("%s: not a TensorBundle index (bad table magic)" % index_path)
    pos = 0
    _mo, pos = _get_varint(footer, pos)
    _ms, pos = _get_varint(footer, pos)
    io, pos = _get_varint(footer, pos)
    isz, pos = _get_varint(footer, pos)
    header, entries = None, {}
    for _, handle in _block_entries(_read_block(f, io, isz, verify)):
        bo, p = _get_varint(handle, 0)
        bs, p = _get_varint(handle, p)
        for key, val in _block_entries(_read_block(f, bo, bs, verify)):
            if key == b"":
                header = dict(num_shards=1, endianness=0)
                for fld, _, v in _proto_fields(val):
                    if fld == 1:
                        header["num_shards"] = v
                    elif fld == 2:
                        header["endianness"] = v
            else:
                entries[key.decode("utf-8")] = _parse_entry(val)
    if header is None:
        raise CheckpointError("%s: no bundle header entry" % index_path)
    if header["endianness"] != 0:
        raise CheckpointError("big-endian checkpoint")
    return header, entries


def read_tensor_bundle(prefix: str, verify: bool = True, names: Optional[Iterable[str]] = None) -> Dict[str, np.ndarray]:
    """All numeric tensors of the checkpoint `prefix` (`prefix.index` + `prefix.data-*`), keyed by their stored names.  String
    tensors (the serialized object graph) are skipped.  verify: check every block's and every tensor's CRC32C."""
    header, entries = read_index(prefix + ".index", verify)
    want = set(names) if names is not None else None
    shards: Dict[int, np.memmap] = {}
    out: Dict[str, np.ndarray] = {}
    for name, e in entries.items():
        if want is not None and name not in want:
            continue
        if e["dtype"] == _DT_STRING or e["sliced"]:
            continue
        if e["dtype"] == _DT_BFLOAT16:
            np_dt, conv = np.uint16, True
        elif e["dtype"] in _DTYPES:
            np_dt, conv = _DTYPES[e["dtype"]], False
        else:
            raise CheckpointError("%s: unsupported dtype enum %d" % (name, e["dtype"]))
        sid = e["shard_id"]
        if sid not in shards:
            path = "%s.data-%05d-of-%05d" % (prefix, sid, header["num_shards"])
            if not os.path.exists(path):
                raise CheckpointError("missing data shard %s" % path)
            shards[sid] = np.memmap(path, dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        count = int(np.prod(e["shape"])) if e["shape"] else 1
        if raw.size != e["size"] or count * np.dtype(np_dt).itemsize != e["size"]:
            raise CheckpointError("%s: size %d does not match shape %s" % (name, e["size"], e["shape"]))
        if verify and e["crc32c"] is not None and _unmask(e["crc32c"]) != _crc32c_np(np.asarray(raw)):
            raise CheckpointError("%s: tensor checksum mismatch" % name)
        arr = np.frombuffer(raw.tobytes(), dtype=np_dt).reshape(e["shape"])
        if conv:
            arr = (arr.astype(np.uint32) << 16).view(np.float32)
        out[name] = arr
    return out


def write_tensor_bundle(prefix: str, tensors: Dict[str, np.ndarray], block_size: int = 4096, restart_interval: int = 16) -> None:
    """Writes `tensors` in the TensorBundle layout (one shard, no compression, masked CRC32Cs, prefix-compressed sorted keys)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    names = sorted(tensors, key=lambda s: s.encode("utf-8"))
    records: List[Tuple[bytes, bytes]] = [(b"", b"\x08\x01\x1a\x02\x08\x01")]       # num_shards = 1, version.producer = 1
    offset = 0
    with open("%s.data-00000-of-00001" % prefix, "wb") as df:
        for n in names:
            a = np.ascontiguousarray(tensors[n])
            if a.dtype not in _DTYPE_IDS:
                raise CheckpointError("%s: dtype %s cannot be written" % (n, a.dtype))
            raw = a.tobytes()
            df.write(raw)
            records.append((n.encode("utf-8"), _entry_proto(_DTYPE_IDS[a.dtype], a.shape, offset, len(raw), _mask(crc32c(raw)))))
            offset += len(raw)
    out = bytearray()

    def emit_block(entries: List[Tuple[bytes, bytes]]) -> Tuple[int, int]:
        body, restarts, prev = bytearray(), [], b""
        for i, (k, v) in enumerate(entries):
            shared = 0
            if i % restart_interval == 0:
                restarts.append(len(body))
            else:
                while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                    shared += 1
            body += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
            prev = k
        if not restarts:
            restarts = [0]
        for r in restarts:
            body += struct.pack("<I", r)
        body += struct.pack("<I", len(restarts))
        off = len(out)
        out.extend(body)
        out.append(0)
        out.extend(struct.pack("<I", _mask(crc32c(bytes(body) + b"\x00"))))
        return off, len(body)

    index: List[Tuple[bytes, bytes]] = []
    cur: List[Tuple[bytes, bytes]] = []
    cur_bytes = 0
    for k, v in records:
        cur.append((k, v))
        cur_bytes += len(k) + len(v) + 3
        if cur_bytes >= block_size:
            o, s = emit_block(cur)
            index.append((cur[-1][0], _put_varint(o) + _put_varint(s)))      # separator = last key of the block
            cur, cur_bytes = [], 0
    if cur:
        o, s = emit_block(cur)
        index.append((cur[-1][0], _put_varint(o) + _put_varint(s)))
    mo, ms = emit_block([])
    io, isz = emit_block(index)
    footer = _put_varint(mo) + _put_varint(ms) + _put_varint(io) + _put_varint(isz)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


# ------------------------------------------------------------------------------------------------ checkpoint directory
def latest_checkpoint(checkpoint_dir: str) -> Optional[str]:
    """tf.train.latest_checkpoint / CheckpointManager.latest_checkpoint (pipeline.py:42-47): the prefix named by the
    `checkpoint` state file, else the highest-numbered `ckpt-N.index`."""
    state = os.path.join(checkpoint_dir, "checkpoint")
    if os.path.exists(state):
        with open(state) as f:
            m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', f.read(), re.M)
        if m:
            p = m.group(1)
            p = p if os.path.isabs(p) else os.path.join(checkpoint_dir, p)
            if os.path.exists(p + ".index"):
                return p
    best = None
    if os.path.isdir(checkpoint_dir):
        for fn in os.listdir(checkpoint_dir):
            m = re.match(r"^(.*-(\d+))\.index$", fn)
            if m and (best is None or int(m.group(2)) > best[0]):
                best = (int(m.group(2)), os.path.join(checkpoint_dir, m.group(1)))
    return best[1] if best else None


def write_checkpoint_state(checkpoint_dir: str, prefix_basename: str) -> None:
    with open(os.path.join(checkpoint_dir, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (prefix_basename, prefix_basename))


# ------------------------------------------------------------------------------------------------ object graph -> weights
_LWW = re.compile(r"layer_with_weights-(\d+)")


def checkpoint_to_weights(tensors: Dict[str, np.ndarray], expected: Optional[Sequence[Tuple[str, tuple]]] = None,
                          keras_layer_order: Optional[Dict[str, List[str]]] = None, strict: bool = True) -> Dict[str, np.ndarray]:
    """Object-graph tensor names -> the engine's variable paths (the keys `fpnmt_set_weight` takes).

    * `<path>/.ATTRIBUTES/VARIABLE_VALUE` -> `<path>`; optimizer state (`optimizer/...`, `.OPTIMIZER_SLOT`), `save_counter` and
      the object-graph blob are dropped.
    * Variables owned by tf.keras layers that the reference reaches through plain attributes / lists (every Dense,
      LayerNormalization and Embedding of models/transformer.py) already carry their Appendix-B path.
    * A Keras *functional* Model (`feature_extractor.retinanet_model`, `feature_extractor.model`, models/retinanet.py:280-304)
      tracks its layers as `layer_with_weights-N`, N = position among the model's weighted layers.  `keras_layer_order` maps the
      model's path to that ordered list of Keras layer names; the default is the build order of `fpnmt.weights.model_spec`,
      i.e. creation order.  UNVERIFIED against TensorFlow (no TF here): Keras orders `model.layers` by graph depth, which may
      differ from creation order for the FPN branches - pass the list printed by `[l.name for l in model.layers if l.weights]`
      from a TF session if a restored model disagrees; `strict` shape checks catch a wrong order for every pair of layers whose
      shapes differ.
    * expected: [(path, shape), ...] (e.g. `fpnmt.weights.model_spec(backbone, vocab)`); with strict, a missing path or a shape
      mismatch raises, naming the key."""
    out: Dict[str, np.ndarray] = {}
    for name, arr in tensors.items():
        if not name.endswith(_SUFFIX) or "/.OPTIMIZER_SLOT/" in name:
            continue
        path = name[:-len(_SUFFIX)]
        if path.startswith("optimizer/") or path == "save_counter" or path.startswith("_"):
            continue
        m = _LWW.search(path)
        if m:
            model_path = path[:m.start()].rstrip("/")
            order = (keras_layer_order or {}).get(model_path)
            if order is None and expected is not None:
                order = _default_layer_order(expected, model_path)
            idx = int(m.group(1))
            if not order or idx >= len(order):
                if strict:
                    raise CheckpointError("%s: no Keras layer order known for model '%s'" % (name, model_path))
                continue
            path = model_path + "/" + order[idx] + path[m.end():]
        out[path] = np.asarray(arr)
    if expected is not None:
        # `feature_extractor.model` shares every layer with `feature_extractor.retinanet_model`; keep one canonical copy
        exp = {p: tuple(s) for p, s, *_ in expected}
        for p in list(out):
            if p not in exp:
                alt = p.replace("/feature_extractor/model/", "/feature_extractor/retinanet_model/")
                if alt in exp and alt not in out:
                    out[alt] = out.pop(p)
        if strict:
            for p, shp in exp.items():
                if p not in out:
                    raise CheckpointError("checkpoint has no variable for '%s'" % p)
                if tuple(out[p].shape) != shp:
                    raise CheckpointError("'%s': checkpoint shape %s, model shape %s" % (p, tuple(out[p].shape), shp))
        out = {p: out[p] for p in exp if p in out}
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


def _default_layer_order(expected: Sequence[Tuple[str, tuple]], model_path: str) -> List[str]:
    order: List[str] = []
    pre = model_path + "/"
    for p, *_ in expected:
        if p.startswith(pre):
            layer = p[len(pre):].rsplit("/", 1)[0]
            if layer not in order:
                order.append(layer)
    return order


def weights_to_checkpoint(weights: Dict[str, np.ndarray], prefix: str, save_counter: int = 1) -> None:
    """The engine's weight dict as a TensorBundle with object-graph tensor names (what train.py:95-96 `ckpt_manager.save()`
    stores for the model side; no optimizer slots, no `_CHECKPOINTABLE_OBJECT_GRAPH` blob - this library's reader and any
    name-based reader restore it, `tf.train.Checkpoint.restore` would additionally want the blob)."""
    t = {k + _SUFFIX: np.asarray(v, dtype=np.float32) for k, v in weights.items()}
    t["save_counter" + _SUFFIX] = np.asarray(save_counter, dtype=np.int64)
    write_tensor_bundle(prefix, t)
    write_checkpoint_state(os.path.dirname(os.path.abspath(prefix)), os.path.basename(prefix))


def load_checkpoint(path: str, backbone: str, vocab: Optional[int] = None, num_layers: Optional[int] = None,
                    verify: bool = True) -> Dict[str, np.ndarray]:
    """`path` = a checkpoint prefix (`.../ckpt-7`) or a CheckpointManager directory -> weight dict for `fpnmt.Engine`."""
    from . import config as C
    from .weights import model_spec
    prefix = path
    if os.path.isdir(path):
        prefix = latest_checkpoint(path)
        if prefix is None:
            raise CheckpointError("no checkpoint found in %s" % path)
    tensors = read_tensor_bundle(prefix, verify=verify)
    if vocab is None:
        k = "transformer/final_layer/kernel" + _SUFFIX
        if k not in tensors:
            raise CheckpointError("checkpoint has no transformer/final_layer/kernel")
        vocab = int(tensors[k].shape[1])
    spec = model_spec(backbone, vocab, num_layers or C.num_layers)
    return checkpoint_to_weights(tensors, expected=[(p, s) for p, s, *_ in spec])
