"""Image-sharded data parallelism for caption inference.

Each image's caption depends only on that image (`Pipeline.predict` is per image, utils/pipeline.py:93; BatchNorm
in inference mode), so a global batch is split contiguously across the ranks of one node, every rank runs its own
engine on its own GPU with a full weight replica, and the ONLY data-path collective is one all-gather of the
int32 caption ids (+ lengths) per batch: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of `total` items; the first `total % world` ranks get one extra."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun); returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


class Communicator:
    """The library's own NCCL communicator (fpnmt_comm_*, include/fpnmt_dlpack.h): the all-gather of the caption ids without
    torch.distributed on the data path.  Bootstrap like NCCL: rank 0 makes the 128-byte unique id, `exchange(id_bytes)` hands it
    to the other ranks (default: torch.distributed's object broadcast when a process group exists; any out-of-band channel
    works - a file, a socket, MPI)."""

    def __init__(self, world: int, rank: int, device: int, exchange=None):
        import ctypes as C
        from . import _lib
        self.lib = _lib.load()
        self._check = _lib.check
        self.world, self.rank, self.device = world, rank, device
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            _lib.check(self.lib.fpnmt_comm_unique_id(buf))
        uid = bytes(buf)
        if world > 1:
            if exchange is None:
                box = [uid]
                dist.broadcast_object_list(box, src=0)
                uid = box[0]
            else:
                uid = exchange(uid)
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._c = C.c_void_p()
        _lib.check(self.lib.fpnmt_comm_create(world, rank, buf, device, C.byref(self._c)))

    def allgather(self, ids_local: torch.Tensor, lens_local: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        b, t = ids_local.shape
        ids_local, lens_local = ids_local.contiguous(), lens_local.contiguous()
        all_ids = torch.empty((self.world * b, t), dtype=torch.int32, device=ids_local.device)
        all_len = torch.empty((self.world * b,), dtype=torch.int32, device=ids_local.device)
        self._check(self.lib.fpnmt_allgather_ids(self._c, ids_local.data_ptr(), lens_local.data_ptr(), b, t, all_ids.data_ptr(),
                                                 all_len.data_ptr(), torch.cuda.current_stream(ids_local.device).cuda_stream))
        return all_ids, all_len

    def close(self):
        if getattr(self, "_c", None) and self._c.value:
            self.lib.fpnmt_comm_destroy(self._c)
            self._c.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def allgather_captions(ids_local: torch.Tensor, lens_local: torch.Tensor, world: Optional[int] = None,
                       comm: Optional[Communicator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather (B_local,T) int32 ids and (B_local,) int32 lengths from every rank (equal B_local) into
    (world*B_local, T) / (world*B_local,), rank-major — one collective: the library's own NCCL communicator when `comm` is given
    (GPU runs), else torch.distributed (gloo in the CPU tests)."""
    if comm is not None and comm.world > 1:
        return comm.allgather(ids_local, lens_local)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return ids_local, lens_local
    world = world or dist.get_world_size()
    b, t = ids_local.shape
    packed = torch.cat([ids_local.reshape(b, t), lens_local.reshape(b, 1).to(ids_local.dtype)], dim=1).contiguous()
    out = torch.empty((world * b, t + 1), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed)
    return out[:, :t].contiguous(), out[:, t].contiguous()


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    """Max of a python float over ranks (timing: report the slowest rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
