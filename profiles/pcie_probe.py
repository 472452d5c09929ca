import torch, time
x = torch.empty(1<<28, dtype=torch.uint8, pin_memory=True)
d = torch.empty(1<<28, dtype=torch.uint8, device='cuda')
for _ in range(2):
    d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter(); 
for _ in range(5): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); h2d = 5*(1<<28)/(time.perf_counter()-t)/1e9
t=time.perf_counter(); 
for _ in range(5): x.copy_(d, non_blocking=True)
torch.cuda.synchronize(); d2h = 5*(1<<28)/(time.perf_counter()-t)/1e9
print("pinned h2d GB/s", h2d, "d2h GB/s", d2h)
# bidirectional
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
x2 = torch.empty(1<<28, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(1<<28, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): x2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); print("bidir total GB/s", 10*(1<<28)/(time.perf_counter()-t)/1e9)
import subprocess; print(subprocess.run("nvidia-smi -q | grep -A6 'GPU Link Info' | head -12; lscpu | grep -E 'Model name|^CPU\\(s\\)'", shell=True, capture_output=True, text=True).stdout)
