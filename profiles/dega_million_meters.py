import sys
sys.path.insert(0,'/root/repo')
import torch
from data_compressor_b200 import Context, synth
ctx=Context(0)
x=synth.meters_15min(1_000_000,96,device='cuda')
ctx.set_profiling(True)
for _ in range(3):
    enc=ctx.dega_encode(x,block_len=0); k1=ctx.kernel_ms()
    y,n_out,st=ctx.dega_decode(enc); k2=ctx.kernel_ms()
assert torch.equal(y.view(torch.int32),x.view(torch.int32))
print({a:round(b,3) for a,b in k1.items() if b}, {a:round(b,3) for a,b in k2.items() if b}, 'ratio', x.numel()*4/enc.total_bytes)
raw=x.numel()*4
print('encode GB/s', raw/(sum(k1.values())*1e-3)/1e9, 'decode GB/s', raw/(sum(k2.values())*1e-3)/1e9)
