"""Throughput of `decode csv` and `encode csv` through the C-ABI host call against the reference's ReadCSV (oracle/_ref, one core --
the reference's csv decoder is a single loop).  Text: the reference's test file printed as "%.2f\\n", repeated.
Usage: python profiles/csv_codec.py [repeats]        (run on the GPU box; prints one JSON line)"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from data_compressor_b200 import Context  # noqa: E402
from oracle.binding import Ref  # noqa: E402

rep = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
vals = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "testfile_f32.npy"))
one = "".join(f"{v:.2f}\n" for v in vals).encode()
text = np.frombuffer(one * rep, dtype=np.uint8)
n_lines = vals.size * rep
ctx = Context(0)
import torch  # noqa: E402
pin = torch.empty(text.size, dtype=torch.uint8).pin_memory()
pin.numpy()[:] = text
out = torch.empty(n_lines, dtype=torch.float32).pin_memory()
n = C.c_uint64(0)
best = 1e9
for it in range(4):
    t0 = time.perf_counter()
    rc = ctx.lib.dc_csv_decode_host(ctx.h, pin.data_ptr(), text.size, 1, b",", out.data_ptr(), n_lines, C.byref(n))
    dt = time.perf_counter() - t0
    assert rc == 0 and n.value == n_lines, (rc, n.value)
    if it:
        best = min(best, dt)
assert np.array_equal(out.numpy()[: vals.size].view(np.uint32), vals.view(np.uint32))
ref = Ref()
sample = one * max(1, min(rep, 40))
t0 = time.perf_counter()
r = ref.csv_decode(sample.decode(), 1, ",")
dt_ref = time.perf_counter() - t0
assert np.array_equal(r.view(np.uint32), out.numpy()[: r.size].view(np.uint32))
# `encode csv` of the same values back to text
txt = torch.empty(text.size + 64, dtype=torch.uint8).pin_memory()
m = C.c_uint64(0)
best_e = 1e9
for it in range(4):
    t0 = time.perf_counter()
    rc = ctx.lib.dc_csv_encode_host(ctx.h, out.data_ptr(), n_lines, 1, b",", 2, txt.data_ptr(), txt.numel(), C.byref(m))
    dt = time.perf_counter() - t0
    assert rc == 0 and m.value == text.size, (rc, m.value)
    if it:
        best_e = min(best_e, dt)
assert np.array_equal(txt.numpy()[: text.size], text)
t0 = time.perf_counter()
rt = ref.csv_encode(r, 1, ",", 2)
dt_ref_e = time.perf_counter() - t0
assert rt == sample
print(json.dumps({"encode_gpu_call_ms": best_e * 1e3, "encode_gpu_text_GBps": text.size / best_e / 1e9, "encode_gpu_Mvalues_per_s": n_lines / best_e / 1e6,
                  "encode_ref_Mvalues_per_s": r.size / dt_ref_e / 1e6}))
print(json.dumps({"text_MB": text.size / 1e6, "lines": n_lines, "gpu_call_ms": best * 1e3, "gpu_text_GBps": text.size / best / 1e9,
                  "gpu_Mlines_per_s": n_lines / best / 1e6, "ref_sample_MB": len(sample) / 1e6, "ref_text_GBps": len(sample) / dt_ref / 1e9,
                  "ref_Mlines_per_s": r.size / dt_ref / 1e6}))
