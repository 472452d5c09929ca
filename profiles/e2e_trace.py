"""Host-side timeline (DC_TRACE=1) of one encode on a context next to one decode on another (config 2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from data_compressor_b200 import Context, synth

C_, S_ = 4096, 86400
ctx = Context(0)
ctx2 = ctx   # one context: the job engine runs the encode and the decode side by side
L_ = ctx.two_wave_block_len(C_, S_)
x = synth.redd_like(C_, S_, device="cuda")
hx = torch.empty((C_, S_), dtype=torch.float32, pin_memory=True)
hx.copy_(x)
torch.cuda.synchronize()
hxn = hx.numpy()
kw = dict(vs=32, nf=100.0, adaptive=True)
cap = int(1.4156e9 / 2.8 * 1.05)
for _ in range(2):
    enc = ctx.dega_encode_host(hxn, block_len=L_, payload_cap=cap, slot=0, **kw)
    ctx.dega_encode_host(hxn, block_len=L_, payload_cap=cap, slot=1, **kw)
    ctx2.dega_decode_host(enc[0], enc[1], enc[2], C_, S_, L_, **kw)
torch.cuda.synchronize()
mode = sys.argv[1] if len(sys.argv) > 1 else "enc_first"
print("---- " + mode, file=sys.stderr, flush=True)
t0 = time.perf_counter()
if mode == "dec_alone":
    ctx2.dega_decode_host(enc[0], enc[1], enc[2], C_, S_, L_, **kw)
elif mode == "enc_alone":
    ctx.dega_encode_host(hxn, block_len=L_, payload_cap=cap, slot=1, **kw)
elif mode == "enc_first":
    pe = ctx.dega_encode_host(hxn, block_len=L_, payload_cap=cap, slot=1, wait=False, **kw)
    pd = ctx2.dega_decode_host(enc[0], enc[1], enc[2], C_, S_, L_, wait=False, **kw)
    pe.wait()
    pd.wait()
else:
    pd = ctx2.dega_decode_host(enc[0], enc[1], enc[2], C_, S_, L_, wait=False, **kw)
    if mode == "dec_first_delay":
        time.sleep(0.002)
    pe = ctx.dega_encode_host(hxn, block_len=L_, payload_cap=cap, slot=1, wait=False, **kw)
    pe.wait()
    pd.wait()
print("%s: %.2f ms" % (mode, (time.perf_counter() - t0) * 1e3))
