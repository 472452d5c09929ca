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


class PendingSizes:
    """The size exchange in flight: started right after a rank's encode, resolved when the container offsets are needed
    (so the collective and the copy of its 16 bytes per rank overlap whatever the rank does next, e.g. the decode of a
    round-trip benchmark, instead of stalling the step)."""

    def __init__(self, rank, world, gathered, work):
        self.rank, self.world, self.gathered, self.work = rank, world, gathered, work

    def _dict(self, allv) -> dict:
        sizes, counts = allv[:, 0].tolist(), allv[:, 1].tolist()
        return {
            "rank": self.rank, "world": self.world, "sizes": sizes, "counts": counts,
            "byte_base": sum(sizes[:self.rank]), "stream_base": sum(counts[:self.rank]),
            "total_bytes": sum(sizes), "total_streams": sum(counts),
        }

    def result(self) -> dict:
        if self.work is not None:
            self.work.wait()
        return self._dict(self.gathered.view(self.world, 2).cpu())


def resolve_all(pendings) -> list:
    """result() of several exchanges in flight (a batch of jobs whose containers are written together): all collectives
    are waited for first, then the 16 bytes per rank of each come back.  (Stacking them on the device for one copy back was
    tried in the bench: the stacking kernel's first launch costs more than the copies it saves.)"""
    pendings = list(pendings)
    for p in pendings:
        if p.work is not None:
            p.work.wait()
            p.work = None
    return [p.result() for p in pendings]


def exchange_sizes_async(total_bytes: int, n_streams: int, device, group=None) -> PendingSizes:
    """Starts the all-gather of (payload bytes, stream count); PendingSizes.result() gives exchange_sizes' dict."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = torch.tensor([int(total_bytes), int(n_streams)], dtype=torch.int64, device=device)
    if world == 1:
        return PendingSizes(rank, world, mine, None)
    flat = torch.empty(world * 2, dtype=torch.int64, device=device)
    work = dist.all_gather_into_tensor(flat, mine, group=group, async_op=True)
    return PendingSizes(rank, world, flat, work)


def global_byte_offsets(local_byte_off: torch.Tensor, byte_base: int) -> torch.Tensor:
    """Container-wide offsets of this rank's streams: local exclusive scan + the rank's base."""
    return local_byte_off + int(byte_base)
