"""Differential fuzzing of the LZMH batch path against the oracle on the GPU box: blocks of random length and statistics
(runs, short periods, text, 16-bit samples, noise) are encoded in one batch, compared byte for byte with the oracle's
encoder, decoded on the GPU and compared with the oracle's decoder.
usage: python profiles/fuzz_lzmh.py [seconds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from data_compressor_b200 import Context
from oracle import binding as ob

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 90.0
ob.build(ref=False)
oracle = ob.Oracle()
ctx = Context(0)
rng = np.random.default_rng(int(time.time()))
t_end = time.time() + budget
rounds = blocks = 0
while time.time() < t_end:
    block = int(rng.choice([1, 2, 3, 127, 128, 129, 274, 275, 402, 403, 404, 1000, 4096, 20000, 65536]))
    n_blocks = int(rng.integers(1, 40))
    tail = int(rng.integers(0, block)) if rng.random() < 0.5 else 0
    total = n_blocks * block + tail
    kind = int(rng.integers(0, 6))
    if kind == 0:
        d = rng.integers(0, 256, total, dtype=np.uint8)
    elif kind == 1:
        d = rng.integers(0, 3, total, dtype=np.uint8)
    elif kind == 2:
        p = int(rng.integers(1, 200))
        d = np.tile(rng.integers(0, 256, p, dtype=np.uint8), total // p + 1)[:total].copy()
        d[rng.integers(0, total, total // 50 + 1)] ^= 1                                   # a few breaks in the period
    elif kind == 3:
        d = np.frombuffer(("".join(f"{v:.2f}\n" for v in rng.random(total // 4 + 2) * 500))[:total].encode(), dtype=np.uint8).copy()
    elif kind == 4:
        w = np.clip(np.cumsum(rng.integers(-5, 6, total // 2 + 1)) + 3000, 0, 65535).astype("<u2")
        d = w.view(np.uint8)[:total].copy()
    else:
        d = np.repeat(rng.integers(0, 256, total // 37 + 1, dtype=np.uint8), 37)[:total].copy()   # runs
    data = torch.from_numpy(d).cuda()
    payload, bit_len, byte_off, status, tot, in_off = ctx.lzmh_encode(data, block_bytes=block)
    assert not status.any().item()
    out, out_len, st2 = ctx.lzmh_decode(payload, tot, bit_len, byte_off, in_off)
    assert not st2.any().item()
    bl, bo, pay = bit_len.cpu().numpy(), byte_off.cpu().numpy(), payload[:tot].cpu().numpy()
    io, ol, oh = in_off.cpu().numpy(), out_len.cpu().numpy(), out.cpu().numpy()
    for s in range(bl.size):
        blk = d[int(io[s]):int(io[s + 1])]
        rc, want, nb = oracle.run_chain([ob.LZMH_ENC], blk, blk.size * 8)
        assert rc == 0 and nb == bl[s] and pay[int(bo[s]):int(bo[s]) + len(want)].tobytes() == want, (block, kind, s)
        rc, dec, nb2 = oracle.run_chain([ob.LZMH_DEC], want, nb, out_cap=blk.size + 4096)
        assert rc == 0 and nb2 == int(ol[s]) * 8 and oh[int(io[s]):int(io[s]) + nb2 // 8].tobytes() == dec, (block, kind, s, "decode")
        blocks += 1
    rounds += 1
print(f"fuzz ok: {rounds} rounds, {blocks} blocks: encoder bytes and decoder bytes equal to the oracle's")
