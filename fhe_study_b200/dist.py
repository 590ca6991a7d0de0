"""Multi-GPU plumbing of the hot path: the path shards over INDEPENDENT units (polynomials,
accumulators, ciphertexts), so there is no data-path collective.  One process per GPU (torch.distributed);
the only communication is one broadcast per key (bootstrapping / key-switching / relinearisation keys) at
load time, over NCCL (NVLink 5 / NVSwitch) on the GPU box and gloo in the CPU tests, plus an optional
all-gather when a caller wants every rank to see all results."""
from __future__ import annotations

from typing import Callable, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of `total` independent units over `world` ranks; the first total % world ranks get
    one extra unit.  Returns [begin, end)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    base, rem = divmod(int(total), world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_key(key, src: int = 0):
    """Broadcast one key tensor (int64 bit patterns of u64 words) from `src` to every rank, in place."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(key, src=src)
    return key


def run_sharded(units, unit_words: int, fn: Callable, gather: bool = False):
    """Apply `fn` (a batched hot-path call) to this rank's contiguous shard of `units` (a [total, unit_words]
    tensor that every rank holds, or a list-like the caller slices).  With gather=True every rank receives
    the concatenation of all shards' results (ragged shards are padded to the largest one for all_gather)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    total = units.shape[0]
    b, e = shard_range(total, rank, world)
    local = fn(units[b:e].contiguous()) if e > b else units.new_empty((0, unit_words))
    if not gather or world == 1:
        return local
    width = local.shape[1] if local.ndim == 2 and local.shape[0] else None
    sizes = [shard_range(total, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    if width is None:  # an empty shard still has to know the result width
        w = torch.zeros(1, dtype=torch.int64, device=units.device)
    else:
        w = torch.tensor([width], dtype=torch.int64, device=units.device)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    width = int(w.item())
    pad = units.new_zeros((biggest, width))
    if local.shape[0]:
        pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
