"""One DEGA encode + decode of the bench workload (for ncu captures; numbers printed under ncu are not bench values)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from data_compressor_b200 import Context, synth

channels = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
block_len = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = Context(0)
x = synth.redd_like(channels, 86400, device="cuda")
for _ in range(steps):
    enc = ctx.dega_encode(x, block_len=block_len)
    y, n_out, status = ctx.dega_decode(enc)
assert torch.equal(y.view(torch.int32), x.view(torch.int32))
print("ok", enc.n_streams, x.numel() * 4 / enc.total_bytes)
