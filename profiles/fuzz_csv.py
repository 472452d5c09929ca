"""Differential fuzzing of the csv codec against the oracle (glibc's printf / strtof) on the GPU box: random float32 bit
patterns printed with every decimal count, and the printed text (plus exponent forms of it) parsed back.
usage: python profiles/fuzz_csv.py [millions of values]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from data_compressor_b200 import Context
from oracle import binding as ob

n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 2_000_000
ob.build(ref=False)
oracle = ob.Oracle()
ctx = Context(0)
rng = np.random.default_rng()
bits = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
x = bits.view(np.float32)
meter = (rng.integers(0, 10 ** 7, n) / 100).astype(np.float32)
checked = 0
for vals, name in ((x, "random bit patterns"), (meter, "two-decimal readings")):
    for d in range(7):
        text = ctx.csv_encode(vals, 1, ",", d)
        assert text == oracle.csv_encode(vals, 1, ",", d), (name, d)
        # whatever was printed parses back to what strtof says (long digit strings of the huge values take the host route)
        fin = text if name != "random bit patterns" else b"\n".join(t for t in text.split(b"\n")[:200000] if len(t) < 40) + b"\n"
        got, want = ctx.csv_decode(fin), oracle.csv_decode(fin)
        assert got.size == want.size and np.array_equal(got.view(np.uint32)[got == got], want.view(np.uint32)[want == want]), (name, d, "decode")
        checked += vals.size
# exponent forms: "%e"-style text of the readings
etext = "".join(f"{v:.6e}\n" for v in meter[:300000]).encode()
got, want = ctx.csv_decode(etext), oracle.csv_decode(etext)
assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
print(f"fuzz ok: {checked} values printed like glibc prints them and parsed back like strtof parses them; 300000 exponent forms")
