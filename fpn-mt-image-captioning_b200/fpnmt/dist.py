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


def allgather_captions(ids_local: torch.Tensor, lens_local: torch.Tensor, world: Optional[int] = None
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather (B_local,T) int32 ids and (B_local,) int32 lengths from every rank (equal B_local) into
    (world*B_local, T) / (world*B_local,), rank-major — one packed collective."""
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
