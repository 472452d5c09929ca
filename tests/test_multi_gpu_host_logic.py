"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: channel sharding + size exchange + container offsets.
The codec on each shard is played by the oracle here (host logic only; the GPU path is covered by -m gpu)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from data_compressor_b200 import shard


def test_channel_ranges_cover_everything():
    for c in (1, 7, 4096, 4097):
        for w in (1, 2, 3, 8):
            r = [shard.channel_range(c, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == c
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def _worker(rank, world, port, n_channels, block_len, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding
    from data_compressor_b200 import synth
    oracle = binding.Oracle()
    x = synth.redd_like(n_channels, 600, device="cpu").numpy()
    lo, hi = shard.channel_range(n_channels, rank, world)
    payload, sizes = bytearray(), []
    for c in range(lo, hi):
        for b in range(0, 600, block_len):
            rc, p, nb = oracle.dega_encode(x[c, b:b + block_len])
            assert rc == 0
            payload += p
            sizes.append(len(p))
    local_off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64)
    info = shard.exchange_sizes(len(payload), len(sizes), torch.device("cpu"))
    goff = shard.global_byte_offsets(local_off, info["byte_base"])
    q.put((rank, bytes(payload), goff.tolist(), info))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_container_layout_matches_single_rank():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_channels, block_len, world = 5, 256, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_channels, block_len, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank reference layout
    from oracle import binding
    from data_compressor_b200 import synth
    oracle = binding.Oracle()
    x = synth.redd_like(n_channels, 600, device="cpu").numpy()
    whole, offs = bytearray(), [0]
    for c in range(n_channels):
        for b in range(0, 600, block_len):
            p = oracle.dega_encode(x[c, b:b + block_len])[1]
            whole += p
            offs.append(len(whole))
    container = bytearray(len(whole))
    all_offs = []
    for rank, payload, goff, info in res:
        assert info["total_bytes"] == len(whole) and info["total_streams"] == len(offs) - 1
        container[info["byte_base"]:info["byte_base"] + len(payload)] = payload
        all_offs += goff[:-1]
    assert bytes(container) == bytes(whole)
    assert all_offs + [len(whole)] == offs
