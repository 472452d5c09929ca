"""Multi-GPU plumbing: shard streams by channel range, exchange per-shard payload sizes, derive container offsets.

The hot path shards by meter channel with no data-path collective: GPU g of G owns channels
[g*C/G, (g+1)*C/G) and encodes them on its own.  The only exchange is one all-gather of per-rank payload byte
totals (8 bytes per rank) followed by a local exclusive scan, which gives every rank its base offset in the final
container.  Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def channel_range(n_channels: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced channel range of `rank` (first n_channels % world ranks get one extra channel)."""
    base, extra = divmod(n_channels, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def exchange_sizes(total_bytes: int, n_streams: int, device, group=None) -> dict:
    """All-gather (payload bytes, stream count) of every rank; returns this rank's base offsets and the totals."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = torch.tensor([int(total_bytes), int(n_streams)], dtype=torch.int64, device=device)
    if world > 1:
        flat = torch.empty(world * 2, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(flat, mine, group=group)
        allv = flat.view(world, 2)
    else:
        allv = mine.view(1, 2)
    sizes = allv[:, 0].cpu().tolist()
    counts = allv[:, 1].cpu().tolist()
    return {
        "rank": rank, "world": world, "sizes": sizes, "counts": counts,
        "byte_base": sum(sizes[:rank]), "stream_base": sum(counts[:rank]),
        "total_bytes": sum(sizes), "total_streams": sum(counts),
    }


def global_byte_offsets(local_byte_off: torch.Tensor, byte_base: int) -> torch.Tensor:
    """Container-wide offsets of this rank's streams: local exclusive scan + the rank's base."""
    return local_byte_off + int(byte_base)
