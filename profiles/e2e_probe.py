import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from data_compressor_b200 import Context, synth
ctx = Context(0)
C_, S_, L_ = 4096, 86400, 1168
x = synth.redd_like(C_, S_, device="cuda")
hx = torch.empty((C_, S_), dtype=torch.float32, pin_memory=True); hx.copy_(x); torch.cuda.synchronize()
hxn = hx.numpy()
for it in range(4):
    t0 = time.perf_counter()
    pay, bl, bo, st = ctx.dega_encode_host(hxn, block_len=L_)
    t1 = time.perf_counter()
    y, n_out, st2 = ctx.dega_decode_host(pay, bl, bo, C_, S_, L_)
    t2 = time.perf_counter()
    print(f"iter {it}: encode_host {1e3*(t1-t0):.1f} ms, decode_host {1e3*(t2-t1):.1f} ms, payload {pay.size/1e6:.1f} MB")
print((y.view(np.uint32) == hxn.view(np.uint32)).all())
