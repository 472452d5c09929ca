"""N-rank NCCL test of the multi-GPU path on real GPUs: every rank encodes its own channel range
on its own GPU, the size exchange (the path's only collective) gives the container offsets, all ranks write ONE container
file with dc_container_write_shard, and every rank decodes its slice of that file back.  Skipped on a box with one GPU;
the host logic alone is covered on CPU by test_multi_gpu_host_logic.py (gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N_CHANNELS, CHANNEL_LEN, BLOCK_LEN = 203, 5000, 512


def _worker(rank, world, port, path, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from data_compressor_b200 import Context, container as ct, shard, synth
        from data_compressor_b200.codec import Encoded
        ctx = Context(rank)
        x_all = synth.redd_like(N_CHANNELS, CHANNEL_LEN, device="cpu", seed=0xC0FFEE)      # same bytes on every rank
        lo, hi = shard.channel_range(N_CHANNELS, rank, world)
        x = x_all[lo:hi].contiguous().to(dev)
        enc = ctx.dega_encode(x, block_len=BLOCK_LEN)
        assert not enc.status.any().item()
        # both forms of the exchange give the same offsets
        sizes = shard.exchange_sizes(enc.total_bytes, enc.n_streams, dev)
        assert shard.exchange_sizes_async(enc.total_bytes, enc.n_streams, dev).result() == sizes
        h, bit_len, payload = ct.prepare_encoded(path, enc, ct.CHAIN_DEGA, 32, 100.0, True, sizes)    # rank 0 creates the file
        dist.barrier()
        ct.write_encoded_shard(path, h, sizes, bit_len, payload)
        dist.barrier()
        # every rank reads ITS slice of the common file and decodes it on its GPU
        g = ct.read_header(path)
        bl = ct.read_index(path, g, sizes["stream_base"], enc.n_streams)
        off = ct.byte_offsets(bl)
        pay = ct.read_payload(path, g, sizes["byte_base"], int(off[-1]))
        d_pay = torch.empty(pay.size + 16, dtype=torch.uint8, device=dev)
        d_pay[: pay.size] = torch.from_numpy(pay).to(dev)
        mine = Encoded(d_pay, torch.from_numpy(bl.astype(np.int64)).to(dev), torch.from_numpy(off.astype(np.int64)).to(dev),
                       torch.zeros(bl.size, dtype=torch.int32, device=dev), int(off[-1]), hi - lo, CHANNEL_LEN, BLOCK_LEN)
        y, n_out, status = ctx.dega_decode(mine)
        ok = (not status.any().item()) and torch.equal(y.view(torch.int32), x.view(torch.int32))
        # and the slice of the NEXT rank (wrap around): the file is one container, not per-rank pieces
        nxt = (rank + 1) % world
        lo2, hi2 = shard.channel_range(N_CHANNELS, nxt, world)
        per = -(-CHANNEL_LEN // BLOCK_LEN)
        all_bits = ct.read_index(path, g)
        all_off = ct.byte_offsets(all_bits)
        s0, s1 = lo2 * per, hi2 * per
        pay2 = ct.read_payload(path, g, int(all_off[s0]), int(all_off[s1] - all_off[s0]))
        d_pay2 = torch.empty(pay2.size + 16, dtype=torch.uint8, device=dev)
        d_pay2[: pay2.size] = torch.from_numpy(pay2).to(dev)
        off2 = (all_off[s0:s1 + 1] - all_off[s0]).astype(np.int64)
        other = Encoded(d_pay2, torch.from_numpy(all_bits[s0:s1].astype(np.int64)).to(dev), torch.from_numpy(off2).to(dev),
                        torch.zeros(s1 - s0, dtype=torch.int32, device=dev), int(off2[-1]), hi2 - lo2, CHANNEL_LEN, BLOCK_LEN)
        y2, _, st2 = ctx.dega_decode(other)
        ok = ok and (not st2.any().item()) and torch.equal(y2.view(torch.int32).cpu(), x_all[lo2:hi2].view(torch.int32))
        q.put((rank, bool(ok), sizes, int(g.n_streams), int(g.payload_bytes), ""))
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()
    except Exception as ex:   # report instead of hanging the parent on q.get
        q.put((rank, False, None, 0, 0, repr(ex)))
        raise


def test_n_rank_container_written_from_gpu_encodes_decodes_back(tmp_path, oracle):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs (run with gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    path = str(tmp_path / "nccl.dcb")
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [r[5] for r in res]
    per = -(-CHANNEL_LEN // BLOCK_LEN)
    assert all(r[1] for r in res), [r[5] for r in res]
    assert all(r[3] == N_CHANNELS * per for r in res)
    assert sum(res[0][2]["sizes"]) == res[0][4] and res[0][2]["byte_base"] == 0
    # the common file holds, stream by stream, what the oracle (= the reference's DCCLI) writes
    from data_compressor_b200 import container as ct, synth
    x_all = synth.redd_like(N_CHANNELS, CHANNEL_LEN, device="cpu", seed=0xC0FFEE).numpy()
    rng = np.random.default_rng(8)
    for st in list(rng.integers(0, N_CHANNELS * per, 16)) + [0, N_CHANNELS * per - 1]:
        c, b = divmod(int(st), per)
        rc, want, nb = oracle.dega_encode(x_all[c, b * BLOCK_LEN:(b + 1) * BLOCK_LEN])
        data, bits = ct.stream_file(path, int(st))
        assert rc == 0 and bits == nb and data == want, st
