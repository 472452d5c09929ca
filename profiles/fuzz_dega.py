"""Differential fuzzing of the DEGA batch path against the oracle on the GPU box: sample distributions the unit tests do
not draw (long constant runs, saw-teeth, heavy tails, near-limit magnitudes, random value sizes and block lengths, static
and adaptive model).  Every stream is round-tripped on the GPU; a random subset is compared byte for byte with the oracle.
usage: python profiles/fuzz_dega.py [seconds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from data_compressor_b200 import Context
from oracle import binding as ob

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
ob.build(ref=False)
oracle = ob.Oracle()
ctx = Context(0)
rng = np.random.default_rng(int(time.time()))
t_end = time.time() + budget
rounds = streams = compared = 0
hundred = np.float32(100.0)
while time.time() < t_end:
    vs = int(rng.choice([17, 20, 24, 32, 33, 48, 64]))
    adaptive = bool(rng.integers(0, 2))
    n_ch = int(rng.integers(1, 400))
    n = int(rng.integers(1, 6000))
    bl = int(rng.choice([0, 0, 96, 777, 1024, 1168, 4096]))
    lim = min(2 ** (vs - 1) - 2, 2 ** 21 - 1)                      # centi-units whose float32 round trip through normalize is exact
    kind = int(rng.integers(0, 7))
    if kind == 0:
        k = rng.integers(0, lim + 1, (n_ch, n))
    elif kind == 1:
        k = np.repeat(rng.integers(0, lim + 1, (n_ch, 1)), n, axis=1)                       # constant channels
    elif kind == 2:
        k = (np.arange(n)[None, :] * rng.integers(1, 50, (n_ch, 1))) % (lim + 1)            # saw-teeth
    elif kind == 3:
        k = np.clip(np.cumsum(rng.integers(-3, 4, (n_ch, n)), axis=1) + lim // 2, 0, lim)   # random walks: tiny differences
    elif kind == 4:
        k = np.where(rng.random((n_ch, n)) < 0.01, lim, 0)                                  # spikes to the limit
    elif kind == 5:
        k = np.clip((rng.pareto(1.2, (n_ch, n)) * 50).astype(np.int64), 0, lim)             # heavy tail
    else:
        k = rng.integers(0, 2, (n_ch, n)) * rng.integers(0, lim + 1, (n_ch, 1))             # two-level
    x = (k.astype(np.float32) / hundred).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    enc = ctx.dega_encode(xd, block_len=bl, vs=vs, nf=100.0, adaptive=adaptive)
    st = enc.status.cpu().numpy()
    y, n_out, st2 = ctx.dega_decode(enc, vs=vs, nf=100.0, adaptive=adaptive)
    ok_streams = st == 0
    per = 1 if bl == 0 else (n + bl - 1) // bl
    L = n if bl == 0 else bl
    # round trip of every stream the encoder accepted (values are exact two-decimal numbers in range: all should be)
    assert ok_streams.all(), (vs, adaptive, kind, int((~ok_streams).sum()))
    assert not st2.any().item() and torch.equal(y.view(torch.int32), xd.view(torch.int32)), (vs, adaptive, kind, bl)
    for s in rng.choice(enc.n_streams, size=min(enc.n_streams, 12), replace=False):
        c, b = divmod(int(s), per)
        rc, want, nb = oracle.run_chain(ob.DEGA_ENC, x[c, b * L:(b + 1) * L], None, vs, 100.0, int(adaptive))
        assert rc == 0 and nb == int(enc.bit_len[s]) and enc.stream_bytes(int(s)) == want, (vs, adaptive, kind, bl, int(s))
        compared += 1
    rounds += 1
    streams += enc.n_streams
print(f"fuzz ok: {rounds} rounds, {streams} streams round-tripped, {compared} compared with the oracle byte for byte")
