"""One LZMH encode + decode of the bench workload (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from data_compressor_b200 import Context, synth

n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
block = 65536
ctx = Context(0)
chans = (n_blocks * block + 2 * 86400 - 1) // (2 * 86400)
data = synth.watts_u16(chans, 86400, device="cuda").reshape(-1)[: n_blocks * block].contiguous()
payload, bit_len, byte_off, status, total, in_off = ctx.lzmh_encode(data, block_bytes=block)
out, out_len, st2 = ctx.lzmh_decode(payload, total, bit_len, byte_off, in_off)
print("ok", n_blocks, data.numel() / total)
