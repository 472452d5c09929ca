import sys, os
sys.path.insert(0, '/root/repo')
import torch
from data_compressor_b200 import Context, synth
ctx = Context(0)
x = synth.redd_like(4096, 86400, device="cuda")
ctx.set_profiling(True)
for bl in (1168, 1024, 1440, 2048, 4096, 0):
    for _ in range(3):
        enc = ctx.dega_encode(x, block_len=bl); km = ctx.kernel_ms()
        y, n_out, status = ctx.dega_decode(enc); km2 = ctx.kernel_ms()
    assert torch.equal(y.view(torch.int32), x.view(torch.int32))
    k = {**{a: b for a, b in km.items() if b}, **{a: b for a, b in km2.items() if b}}
    print(bl, enc.n_streams, round(x.numel() * 4 / enc.total_bytes, 4), {a: round(b, 3) for a, b in k.items()}, round(sum(k.values()), 2))
