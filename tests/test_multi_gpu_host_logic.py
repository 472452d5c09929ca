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


def _worker(rank, world, port, n_channels, block_len, q, container_path=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding
    from data_compressor_b200 import synth
    oracle = binding.Oracle()
    x = synth.redd_like(n_channels, 600, device="cpu").numpy()
    lo, hi = shard.channel_range(n_channels, rank, world)
    payload, sizes, bits = bytearray(), [], []
    for c in range(lo, hi):
        for b in range(0, 600, block_len):
            rc, p, nb = oracle.dega_encode(x[c, b:b + block_len])
            assert rc == 0
            payload += p
            sizes.append(len(p))
            bits.append(nb)
    local_off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64)
    info = shard.exchange_sizes(len(payload), len(sizes), torch.device("cpu"))
    goff = shard.global_byte_offsets(local_off, info["byte_base"])
    # the deferred forms give the same dictionaries: one at a time, and several jobs resolved with one copy back
    pend = [shard.exchange_sizes_async(len(payload) + k, len(sizes), torch.device("cpu")) for k in range(3)]
    assert pend[0].result() == info
    many = shard.resolve_all(pend)
    assert many[0] == info and [m["sizes"][rank] for m in many] == [len(payload) + k for k in range(3)]
    if container_path is not None:
        # every rank writes its own slice of ONE file: rank 0 creates it, a barrier, then position-based writes
        from data_compressor_b200 import container as ct
        bpc = (600 + block_len - 1) // block_len
        h = ct.make_header(ct.CHAIN_DEGA, info["total_streams"], info["total_bytes"], info["total_streams"] // bpc, 600, block_len)
        if rank == 0:
            ct.create(container_path, h)
        dist.barrier()
        ct.write_shard(container_path, h, info["stream_base"], np.array(bits, dtype=np.uint64), info["byte_base"],
                       np.frombuffer(bytes(payload), dtype=np.uint8))
    q.put((rank, bytes(payload), goff.tolist(), info))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_container_layout_matches_single_rank(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_channels, block_len, world = 5, 256, 2
    path = str(tmp_path / "two_ranks.dcb")
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_channels, block_len, q, path)) for r in range(world)]
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
    # the file the two ranks wrote together holds the single-rank layout, stream by stream
    from data_compressor_b200 import container as ct
    h = ct.read_header(path)
    assert (h.n_channels, h.n_streams, h.payload_bytes) == (n_channels, len(offs) - 1, len(whole))
    assert ct.read_payload(path).tobytes() == bytes(whole)
    assert ct.byte_offsets(ct.read_index(path)).tolist() == offs
    rc, y = oracle.dega_decode(*ct.stream_file(path, 7))
    assert rc == 0 and np.array_equal(y.view(np.uint32), x[7 // 3, (7 % 3) * block_len:(7 % 3 + 1) * block_len].view(np.uint32))
