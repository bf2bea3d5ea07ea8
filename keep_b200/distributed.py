"""Multi-GPU plumbing: one process per GPU, tiles sharded contiguously, one all-gather of the results.

The reference has no multi-GPU inference (every WSI script pins `device='cuda:0'`,
WSI_evaluation/zeroshot_detection_WSI.py:21). Tiles are independent through encode_image and the
similarity rows are independent, so the path shards embarrassingly: rank r of R encodes the
contiguous tile range [r*ceil(N/R), min(N, (r+1)*ceil(N/R))), weights are replicated, and the only exchange
is ONE all-gather (NCCL over NVLink/NVSwitch; gloo on CPU in the tests) of the per-rank [N/R, D] embeddings —
or of the [N/R, P] probabilities when only those are needed.  76.8 MB for 50k x 768 fp16 rows, against
~0.5 s of per-rank compute, so nothing is gained by overlapping it.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT). Returns
    (rank, local_rank, world_size); a no-op single-process group is NOT created when WORLD_SIZE is 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_size(n: int, world_size: int) -> int:
    return (n + world_size - 1) // world_size


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard of `n` rows owned by `rank`; trailing ranks may own fewer rows or none."""
    per = shard_size(n, world_size)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate the per-rank row shards (as produced by `shard_range`) into the full [n_total, ...] tensor on
    every rank. Shards are padded to the common length ceil(n_total/R) for the single fixed-size all-gather and
    the padding is dropped afterwards."""
    rank, ws = world()
    if ws == 1:
        assert local.shape[0] == n_total
        return local
    per = shard_size(n_total, ws)
    lo, hi = shard_range(n_total, rank, ws)
    assert local.shape[0] == hi - lo, f"rank {rank}: shard has {local.shape[0]} rows, expected {hi - lo}"
    if local.shape[0] < per:
        pad = torch.zeros((per - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = torch.empty((ws * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:n_total]


def encode_tiles_sharded(model, n_total: int, load_tiles: Callable[[int, int], torch.Tensor],
                         batch: int = 1024, gather_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Every rank encodes its shard of the `n_total` tiles in streamed batches of `batch`
    (`load_tiles(lo, hi)` returns tiles [hi-lo, 3, 224, 224] on the model's device) and the embeddings are
    all-gathered: returns [n_total, D] on every rank."""
    rank, ws = world()
    lo, hi = shard_range(n_total, rank, ws)
    outs = []
    for b0 in range(lo, hi, batch):
        b1 = min(hi, b0 + batch)
        outs.append(model.encode_image(load_tiles(b0, b1)))
    dev = model.logit_scale.device
    D = model.config.projection_dim
    local = torch.cat(outs, dim=0) if outs else torch.empty(0, D, dtype=torch.float32, device=dev)
    if gather_dtype is not None:
        local = local.to(gather_dtype)
    full = all_gather_rows(local, n_total)
    return full.float() if gather_dtype is not None else full


def barrier_max_ms(ms_local: float, device) -> float:
    """Device-timed duration, max over ranks (the number a multi-GPU measurement must report)."""
    rank, ws = world()
    if ws == 1:
        return ms_local
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
