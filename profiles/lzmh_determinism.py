"""Race check of the warp-cooperative LZMH kernels: the same 4096 blocks encoded and decoded ten times must give the same
bytes every time (and the decoded bytes must be the input wherever the reference restores it)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from data_compressor_b200 import Context, synth

n_blocks, block = 4096, 65536
ctx = Context(0)
chans = (n_blocks * block + 2 * 86400 - 1) // (2 * 86400)
data = synth.watts_u16(chans, 86400, device="cuda").reshape(-1)[: n_blocks * block].contiguous()
ref_pay = ref_out = None
for it in range(10):
    payload, bit_len, byte_off, status, total, in_off = ctx.lzmh_encode(data, block_bytes=block)
    out, out_len, st2 = ctx.lzmh_decode(payload, total, bit_len, byte_off, in_off)
    assert not status.any().item() and not st2.any().item()
    pay = payload[:total].clone()
    if ref_pay is None:
        ref_pay, ref_out, ref_len, ref_bits = pay, out.clone(), out_len.clone(), bit_len.clone()
    else:
        assert torch.equal(pay, ref_pay) and torch.equal(bit_len, ref_bits), f"encode differs in repetition {it}"
        assert torch.equal(out_len, ref_len), f"decode lengths differ in repetition {it}"
        idx = torch.arange(block, device="cuda").view(1, block)
        valid = idx < out_len.view(-1, 1).to(torch.int64)              # bytes past a block's decoded length are never written
        a, b = out[: n_blocks * block].view(n_blocks, block), ref_out[: n_blocks * block].view(n_blocks, block)
        assert torch.equal(a[valid], b[valid]), f"decode bytes differ in repetition {it}"
full = out_len == (in_off[1:] - in_off[:-1])
o2, d2 = out[: n_blocks * block].view(n_blocks, block), data.view(n_blocks, block)
assert torch.equal(o2[full], d2[full])
print("deterministic over 10 repetitions;", int(full.sum().item()), "of", n_blocks, "blocks restored in full, the rest lose the reference's trailing zero byte")
