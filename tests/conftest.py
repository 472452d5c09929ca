import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


# ---------------------------------------------------------------------------------------------------------
# golden vectors (tests/golden/golden.npz, produced from the real reference by tests/golden/make_golden.py)
class Golden:
    def __init__(self):
        here = os.path.join(ROOT, "tests", "golden")
        self.z = np.load(os.path.join(here, "golden.npz"))
        self.testfile = np.load(os.path.join(here, "testfile_f32.npy"))
        self.names = [str(n) for n in self.z["names"]]

    def get(self, name):
        z = self.z
        in_bits, vs, adaptive, rc, out_bits = (int(v) for v in z[f"{name}/meta"])
        inp = z[f"{name}/input"]
        if inp.size == 0 and in_bits == self.testfile.size * 32:
            inp = self.testfile.view(np.uint8)
        return dict(name=name, input=inp, in_bits=in_bits, chain=[int(c) for c in z[f"{name}/chain"]], vs=vs,
                    nf=float(z[f"{name}/nf"][0]), adaptive=adaptive, rc=rc, out_bits=out_bits, out=z[f"{name}/out"],
                    sha256=bytes(z[f"{name}/sha256"]).decode())


@pytest.fixture(scope="session")
def golden():
    return Golden()


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build(ref=False)
    return binding.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The real reference (oracle/_ref).  Present in the build container; on the GPU box only if it travelled."""
    from oracle import binding
    if os.path.isdir("/root/reference/DataCompressor"):
        binding.build(ref=True)
    if not binding.Ref.available():
        pytest.skip("oracle/_ref/libdcref.so not available")
    return binding.Ref()


# ---------------------------------------------------------------------------------------------------------
# host-compiled device bodies (tests/host_sim): logic tests without a GPU
class Sim:
    def __init__(self):
        d = os.path.join(ROOT, "tests", "host_sim")
        so = os.path.join(d, "_build", "libdcsim.so")
        os.makedirs(os.path.dirname(so), exist_ok=True)
        srcs = [os.path.join(d, "sim.cpp"), os.path.join(d, "host_shim.h")] + \
            [os.path.join(ROOT, "data_compressor_b200", "csrc", f) for f in ("bitio.cuh", "bac.cuh", "dega_device.cuh", "kernels.h",
                                                                             "lzmh_device.cuh", "csv_device.cuh")]
        srcs = [s for s in srcs if os.path.exists(s)]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
            subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
                            "-Wno-maybe-uninitialized", "-I", cuda_inc, "-o", so, os.path.join(d, "sim.cpp")], check=True)
        L = C.CDLL(so)
        L.sim_stage_run.restype = C.c_int
        L.sim_stage_run.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint32, C.c_float, C.c_int, C.c_void_p, C.c_uint64,
                                    C.POINTER(C.c_uint64)]
        L.sim_bac_redo_count.restype = C.c_uint64
        L.sim_bac_redo_count.argtypes = [C.c_int]
        L.sim_check_div_table.restype = C.c_uint64
        L.sim_check_div_table.argtypes = [C.c_uint32]
        L.sim_dega_encode_batch.restype = C.c_int
        L.sim_dega_encode_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_float, C.c_int, C.c_void_p,
                                            C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.sim_dega_decode_batch.restype = C.c_int
        L.sim_dega_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_float,
                                            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sim_check_narrow_w.restype = C.c_uint64
        L.sim_check_narrow_w.argtypes = [C.c_uint64, C.c_uint64]
        L.sim_dense_reader.restype = C.c_uint32
        L.sim_dense_reader.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        L.sim_carry_writer.restype = C.c_uint64
        L.sim_carry_writer.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
        L.sim_csv_field.restype = C.c_float
        L.sim_csv_field.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_int)]
        L.sim_csv_encode.restype = C.c_uint64
        L.sim_csv_encode.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_char, C.c_uint32, C.c_void_p, C.c_uint64]
        self.lib = L

    def csv_field(self, field: str):
        """(value, hard) of csv_to_float on one field."""
        b = field.encode()
        hard = C.c_int(0)
        v = self.lib.sim_csv_field(b, len(b), C.byref(hard))
        return np.float32(v), bool(hard.value)

    def csv_encode(self, values, column=1, separator=",", decimals=2) -> bytes:
        x = np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
        out = np.empty(x.size * (column + 56) + 16, dtype=np.uint8)
        n = int(self.lib.sim_csv_encode(x.ctypes.data if x.size else None, x.size, column, separator.encode()[:1], decimals, out.ctypes.data, out.size))
        assert n <= out.size
        return out[:n].tobytes()

    def stage_run(self, stage, data, nbits=None, vs=32, nf=100.0, adaptive=0, cap=None):
        src = np.frombuffer(bytes(data), dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) \
            else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        if nbits is None:
            nbits = src.size * 8
        cap = cap or max(1 << 16, src.size * 40 + 4096)
        out = np.zeros(cap, dtype=np.uint8)
        ob = C.c_uint64(0)
        rc = self.lib.sim_stage_run(stage, src.ctypes.data if src.size else None, nbits, vs, nf, int(adaptive), out.ctypes.data, cap,
                                    C.byref(ob))
        nb = int(ob.value) if rc == 0 else 0
        return rc, out[: (nb + 7) // 8].tobytes(), nb

    def dega_encode_batch(self, x, block_len, vs=32, nf=100.0, adaptive=1):
        x = np.ascontiguousarray(x, dtype=np.float32)
        nc, cl = x.shape
        bl = block_len or max(cl, 1)
        n = nc * (1 if cl == 0 else (cl + bl - 1) // bl)
        cap = x.size * 9 + n * 16 + 64
        payload = np.zeros(cap, np.uint8)
        bit_len = np.zeros(n, np.uint64)
        byte_off = np.zeros(n + 1, np.uint64)
        status = np.zeros(n, np.int32)
        total = C.c_uint64(0)
        rc = self.lib.sim_dega_encode_batch(x.ctypes.data, nc, cl, block_len, vs, nf, int(adaptive), payload.ctypes.data, cap,
                                            bit_len.ctypes.data, byte_off.ctypes.data, status.ctypes.data, C.byref(total))
        return rc, payload[: total.value], bit_len, byte_off, status

    def dega_decode_batch(self, payload, bit_len, byte_off, nc, cl, block_len, vs=32, nf=100.0, adaptive=1):
        payload = np.ascontiguousarray(payload, np.uint8)
        n = bit_len.size
        out = np.zeros((nc, cl), np.float32)
        n_out = np.zeros(n, np.uint64)
        status = np.zeros(n, np.int32)
        rc = self.lib.sim_dega_decode_batch(payload.ctypes.data, bit_len.ctypes.data, byte_off.ctypes.data, nc, cl, block_len, vs, nf,
                                            int(adaptive), out.ctypes.data, n_out.ctypes.data, status.ctypes.data)
        return rc, out, n_out, status


@pytest.fixture(scope="session")
def sim():
    return Sim()


# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="session")
def ctx():
    """A live dc_ctx on cuda:0 (GPU tests only).  Fails -- never skips -- when the extension is missing."""
    import torch
    assert torch.cuda.is_available(), "GPU test running without a GPU"
    from data_compressor_b200 import Context
    c = Context(0)
    yield c
    c.close()


def split_streams(x, block_len):
    """(channel, block) streams of a [n_channels, channel_len] array, in container order."""
    nc, cl = x.shape
    bl = block_len or max(cl, 1)
    out = []
    for c in range(nc):
        for b in range(0, max(cl, 1), bl):
            out.append(x[c, b:b + bl])
    return out


# ---------------------------------------------------------------------------------------------------------
def straddling_bits(n_straddle, tail, adaptive=1, prefix=(), resolve=None):
    """Input bits for the arithmetic coder that keep the coding interval straddling the midpoint for `n_straddle`
    symbols (pure-Python model of bac.c:54-81,107-137): every such symbol adds pending bits (E3 steps), i.e. a
    long run of output bits that a later symbol resolves with or without a carry.  resolve="up"/"down" appends the
    symbol that puts the whole interval above/below the midpoint (a carry into the run / none); `tail` follows."""
    freq, cum = [0, 1, 1, 1], [3, 2, 1, 0]
    s2i, i2s = [1, 2], [0, 0, 1, 0]
    lo, hi = 0, 0xFFFF
    out = []

    def code(bit):
        nonlocal lo, hi
        idx = s2i[bit]
        rng = hi - lo + 1
        hi = lo + (rng * cum[idx - 1]) // cum[0] - 1
        lo = lo + (rng * cum[idx]) // cum[0]
        while True:
            if hi < 0x8000:
                pass
            elif lo >= 0x8000:
                lo -= 0x8000
                hi -= 0x8000
            elif lo >= 0x4000 and hi < 0xC000:
                lo -= 0x4000
                hi -= 0x4000
            else:
                break
            lo, hi = 2 * lo, 2 * hi + 1
        if adaptive:
            if cum[0] == 16383:
                c = 0
                for k in (3, 2, 1, 0):
                    freq[k] = (freq[k] + 1) // 2
                    cum[k] = c
                    c += freq[k]
            k = idx
            while freq[k] == freq[k - 1]:
                k -= 1
            if k < idx:
                sk, si = i2s[k], i2s[idx]
                i2s[k], i2s[idx] = si, sk
                s2i[sk], s2i[si] = idx, k
            freq[k] += 1
            while k > 0:
                k -= 1
                cum[k] += 1
        out.append(bit)

    for b in prefix:
        code(b)
    for _ in range(n_straddle):
        rng = hi - lo + 1
        # the symbol whose sub-interval contains the midpoint 0x8000 (falls back to bit 0 if that is EOF's)
        pick = 0
        for bit in (0, 1):
            idx = s2i[bit]
            h = lo + (rng * cum[idx - 1]) // cum[0] - 1
            l = lo + (rng * cum[idx]) // cum[0]
            if l < 0x8000 <= h:
                pick = bit
        code(pick)
    for _ in range(64 if resolve else 0):
        rng = hi - lo + 1
        a = lo + (rng * cum[1]) // cum[0]          # index 1 (the more frequent symbol) owns [a, hi], index 2 [.., a-1]
        if lo < 0x8000 <= hi and resolve == "up" and a > 0x8000 and s2i[0] != s2i[1]:
            code(i2s[1])                           # midpoint inside index 2's part: index 1 lies above it
            break
        if lo < 0x8000 <= hi and resolve == "down" and a <= 0x8000:
            code(i2s[2])                           # midpoint inside index 1's part: index 2 lies below it
            break
        code(i2s[1] if a <= 0x8000 else i2s[2])    # keep straddling
    for b in tail:
        code(b)
    return np.array(out, dtype=np.uint8)
