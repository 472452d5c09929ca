import sys, os
sys.path.insert(0, os.getcwd())
import torch
from data_compressor_b200 import Context, synth
ctx = Context(0); ctx.set_profiling(True)
x = synth.redd_like(4096, 86400, device="cuda")
for bl in (0, 16384):
    enc = ctx.dega_encode(x, block_len=bl)
    for _ in range(2):
        enc = ctx.dega_encode(x, block_len=bl, out=enc); k1 = ctx.kernel_ms()
        y, n, s = ctx.dega_decode(enc); k2 = ctx.kernel_ms()
    print(bl, {k: round(v, 2) for k, v in k1.items() if v}, {k: round(v, 2) for k, v in k2.items() if v})
