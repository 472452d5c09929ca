"""Where does the time of the host-buffer calls go?  Encode alone, decode alone, and both at once on two contexts
(config 2, pinned host memory).  Run on the GPU box: python profiles/e2e_overlap_probe.py"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
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
    for slot in (0, 1):
        enc = ctx.dega_encode_host(hxn, block_len=L_, payload_cap=cap, slot=slot, **kw)
    ctx2.dega_decode_host(enc[0], enc[1], enc[2], C_, S_, L_, **kw)


enc = ctx.dega_encode_host(hxn, block_len=L_, payload_cap=cap, slot=2, **kw)   # a stable input for the decodes


def t(f, n=3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print("encode alone   %.2f ms" % t(lambda: ctx.dega_encode_host(hxn, block_len=L_, payload_cap=cap, slot=0, **kw)))
print("decode alone   %.2f ms" % t(lambda: ctx2.dega_decode_host(enc[0], enc[1], enc[2], C_, S_, L_, **kw)))


def both():
    pd = ctx2.dega_decode_host(enc[0], enc[1], enc[2], C_, S_, L_, wait=False, **kw)
    pe = ctx.dega_encode_host(hxn, block_len=L_, payload_cap=cap, slot=1, wait=False, **kw)
    t0 = time.perf_counter()
    pe.wait()
    t1 = time.perf_counter()
    pd.wait()
    both.t_enc.append(t1 - t0)


both.t_enc = []
print("both, async    %.2f ms (encode finished after %s ms)" % (t(both), ["%.1f" % (v * 1e3) for v in both.t_enc]))


def both_threads():
    th = threading.Thread(target=lambda: ctx2.dega_decode_host(enc[0], enc[1], enc[2], C_, S_, L_, **kw))
    th.start()
    ctx.dega_encode_host(hxn, block_len=L_, payload_cap=cap, slot=0, **kw)
    th.join()


print("both, threads  %.2f ms" % t(both_threads))
for chunks in ("30311", "37888", "50518", "60621", "75776", "101035", "151552"):
    os.environ["DC_HOST_CHUNK_STREAMS"] = chunks
    print("chunk streams %9s: encode %.2f  decode %.2f  both %.2f ms" % (
        chunks, t(lambda: ctx.dega_encode_host(hxn, block_len=L_, payload_cap=cap, slot=1, **kw)),
        t(lambda: ctx2.dega_decode_host(enc[0], enc[1], enc[2], C_, S_, L_, **kw)), t(both)))
# raw copies, same sizes, two plain streams
d = torch.empty(1415577600, dtype=torch.uint8, device="cuda")
d2 = torch.empty(1415577600, dtype=torch.uint8, device="cuda")
h2 = torch.empty(1415577600, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
hb = hx.view(torch.uint8).reshape(-1)


def h2d():
    with torch.cuda.stream(s1):
        d.copy_(hb, non_blocking=True)
    s1.synchronize()


def d2h():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    s2.synchronize()


def duplex():
    with torch.cuda.stream(s1):
        d.copy_(hb, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    s1.synchronize()
    s2.synchronize()


print("raw 1.4 GB: h2d %.2f  d2h %.2f  duplex %.2f ms" % (t(h2d), t(d2h), t(duplex)))
