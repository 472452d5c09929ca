"""ctypes bindings for the checker libraries (TEST INFRASTRUCTURE ONLY).

* ``Oracle``  -> oracle/_build/libdcoracle.so, our plain-C restatement (oracle/dc_oracle.c)
* ``Ref``     -> oracle/_ref/libdcref.so, the UNMODIFIED reference DCLib/DCIOLib compiled from
                 /root/reference by oracle/Makefile, driven through oracle/ref_driver.c

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  The product package (data_compressor_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libdcoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libdcref.so")
REF_CLI = os.path.join(HERE, "_ref", "DCCLI")

# enum dco_stage_id (oracle/dc_oracle.h)
NORMALIZE, DENORMALIZE, DIFF_ENC, DIFF_DEC, SEG_ENC, SEG_DEC, BAC_ENC, BAC_DEC, LZMH_ENC, LZMH_DEC = range(10)
COPY = 12   # DCO_COPY == DC_STAGE_COPY

#: oracle stage id -> (DCCLI encoder name, encode flag)   (DCLib/src/enc_dec.c:51-60)
STAGE_TO_CLI = {
    NORMALIZE: ("normalize", 1), DENORMALIZE: ("normalize", 0),
    DIFF_ENC: ("diff", 1), DIFF_DEC: ("diff", 0),
    SEG_ENC: ("seg", 1), SEG_DEC: ("seg", 0),
    BAC_ENC: ("bac", 1), BAC_DEC: ("bac", 0),
    LZMH_ENC: ("lzmh", 1), LZMH_DEC: ("lzmh", 0),
    COPY: ("copy", 1),
}

DEGA_ENC = [NORMALIZE, DIFF_ENC, SEG_ENC, BAC_ENC]
DEGA_DEC = [BAC_DEC, SEG_DEC, DIFF_DEC, DENORMALIZE]


def build(ref: bool = True) -> None:
    """Compile the checker libraries (building the checker is not using it)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(buf), dtype=np.uint8)
    a = np.ascontiguousarray(buf)
    return a.view(np.uint8).reshape(-1)


class Oracle:
    """Our C restatement.  All results are (code, bytes, nbits)."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        L = self.lib
        L.dco_run_chain.restype = C.c_int
        L.dco_run_chain.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_uint64, C.c_uint, C.c_float,
                                    C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.dco_aggregate.restype = C.c_uint64
        L.dco_aggregate.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_void_p]
        L.dco_dega_encode_batch.restype = C.c_int64
        L.dco_dega_encode_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint, C.c_float, C.c_int]

    def set_block_size(self, bits: int) -> None:
        """`blocksize` of the copy stages in the chains run from now on (default 8)."""
        self.lib.dco_set_block_size(C.c_uint(int(bits)))

    def run_chain(self, stages, data, nbits=None, vs=32, nf=100.0, adaptive=1, out_cap=None):
        src = _as_u8(data)
        if nbits is None:
            nbits = src.size * 8
        if out_cap is None:
            out_cap = max(4096, src.size * 24 + 4096)
        out = np.zeros(out_cap, dtype=np.uint8)
        out_bits = C.c_uint64(0)
        st = (C.c_int * len(stages))(*stages)
        rc = self.lib.dco_run_chain(st, len(stages), src.ctypes.data if src.size else None, nbits, vs, nf,
                                    int(adaptive), out.ctypes.data, out_cap, C.byref(out_bits))
        nb = int(out_bits.value)
        return rc, out[: (nb + 7) // 8].tobytes(), nb

    def dega_encode(self, samples, vs=32, nf=100.0, adaptive=1):
        x = np.ascontiguousarray(samples, dtype=np.float32)
        return self.run_chain(DEGA_ENC, x, x.size * 32, vs, nf, adaptive)

    def dega_decode(self, payload, nbits=None, vs=32, nf=100.0, adaptive=1, out_cap=None):
        rc, raw, nb = self.run_chain(DEGA_DEC, payload, nbits, vs, nf, adaptive, out_cap=out_cap)
        return rc, np.frombuffer(raw[: (nb // 32) * 4], dtype=np.float32).copy()

    def csv_decode(self, text, column: int = 1, separator: str = ",") -> np.ndarray:
        """`decode csv` (csv.c:13-44) on a text buffer."""
        src = np.frombuffer(text.encode() if isinstance(text, str) else bytes(text), dtype=np.uint8)
        self.lib.dco_csv_decode.restype = C.c_uint64
        self.lib.dco_csv_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_char, C.c_void_p, C.c_uint64]
        out = np.empty(src.size + 1, dtype=np.float32)
        n = int(self.lib.dco_csv_decode(src.ctypes.data if src.size else None, src.size, int(column), separator.encode()[:1],
                                        out.ctypes.data, out.size))
        return out[:n].copy()

    def csv_encode(self, values, column: int = 1, separator: str = ",", decimals: int = 2) -> bytes:
        """`encode csv` (csv.c:46-65): the text of the values."""
        x = np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
        self.lib.dco_csv_encode.restype = C.c_uint64
        self.lib.dco_csv_encode.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_char, C.c_uint, C.c_void_p, C.c_uint64]
        out = np.empty(x.size * (column + 56) + 16, dtype=np.uint8)
        n = int(self.lib.dco_csv_encode(x.ctypes.data if x.size else None, x.size, int(column), separator.encode()[:1], int(decimals),
                                        out.ctypes.data, out.size))
        return out[:n].tobytes()

    def aggregate(self, samples, num_values: int) -> np.ndarray:
        """`encode aggregate num_values=k` on one channel (aggregate.c:9-27)."""
        x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
        out = np.empty((x.size + num_values - 1) // num_values + 1, dtype=np.float32)
        n = int(self.lib.dco_aggregate(x.ctypes.data if x.size else None, x.size, int(num_values), out.ctypes.data))
        return out[:n].copy()

    def dega_encode_batch(self, samples2d, vs=32, nf=100.0, adaptive=1) -> int:
        x = np.ascontiguousarray(samples2d, dtype=np.float32)
        assert x.ndim == 2
        return int(self.lib.dco_dega_encode_batch(x.ctypes.data, x.shape[0], x.shape[1], vs, nf, int(adaptive)))


class Ref:
    """The real reference, in process (not re-entrant: one call at a time per process)."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_run_chain.restype = C.c_longlong
        L.ref_run_chain.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_ulonglong,
                                    C.c_uint, C.c_float, C.c_int, C.c_void_p, C.c_ulonglong,
                                    C.POINTER(C.c_ulonglong), C.POINTER(C.c_int)]

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def csv_decode(self, text, column: int = 1, separator: str = ",") -> np.ndarray:
        """`decode csv` through the reference's own ReadCSV()."""
        if not hasattr(self.lib, "ref_csv_decode"):
            raise RuntimeError("oracle/_ref/libdcref.so predates ref_csv_decode: rebuild it where /root/reference exists")
        self.lib.ref_csv_decode.restype = C.c_longlong
        self.lib.ref_csv_decode.argtypes = [C.c_void_p, C.c_ulonglong, C.c_uint, C.c_char, C.c_void_p, C.c_ulonglong]
        src = np.frombuffer(text.encode() if isinstance(text, str) else bytes(text), dtype=np.uint8)
        out = np.empty(src.size + 1, dtype=np.float32)
        n = int(self.lib.ref_csv_decode(src.ctypes.data if src.size else None, src.size, int(column), separator.encode()[:1],
                                        out.ctypes.data, out.size))
        if n < 0:
            raise RuntimeError(f"reference csv failed: {n}")
        return out[:n].copy()

    def csv_encode(self, values, column: int = 1, separator: str = ",", decimals: int = 2) -> bytes:
        """`encode csv` through the reference's own WriteCSV()."""
        if not hasattr(self.lib, "ref_csv_encode"):
            raise RuntimeError("oracle/_ref/libdcref.so predates ref_csv_encode: rebuild it where /root/reference exists")
        self.lib.ref_csv_encode.restype = C.c_longlong
        self.lib.ref_csv_encode.argtypes = [C.c_void_p, C.c_ulonglong, C.c_uint, C.c_char, C.c_uint, C.c_void_p, C.c_ulonglong]
        x = np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
        out = np.empty(x.size * (column + 56) + 16, dtype=np.uint8)
        n = int(self.lib.ref_csv_encode(x.ctypes.data if x.size else None, x.size, int(column), separator.encode()[:1], int(decimals),
                                        out.ctypes.data, out.size))
        if n < 0:
            raise RuntimeError(f"reference csv failed: {n}")
        return out[:n].tobytes()

    def aggregate(self, samples, num_values: int) -> np.ndarray:
        """`encode aggregate num_values=k` through the reference's own Aggregate()."""
        if not hasattr(self.lib, "ref_aggregate"):
            raise RuntimeError("oracle/_ref/libdcref.so predates ref_aggregate: rebuild it where /root/reference exists")
        self.lib.ref_aggregate.restype = C.c_longlong
        self.lib.ref_aggregate.argtypes = [C.c_void_p, C.c_ulonglong, C.c_uint, C.c_void_p, C.c_ulonglong]
        x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
        out = np.empty((x.size + num_values - 1) // num_values + 1, dtype=np.float32)
        n = int(self.lib.ref_aggregate(x.ctypes.data if x.size else None, x.size, int(num_values), out.ctypes.data, out.size))
        if n < 0:
            raise RuntimeError(f"reference aggregate failed: {n}")
        return out[:n].copy()

    def set_block_size(self, bits: int) -> None:
        if not hasattr(self.lib, "ref_set_block_size"):
            raise RuntimeError("oracle/_ref/libdcref.so predates ref_set_block_size: rebuild it where /root/reference exists")
        self.lib.ref_set_block_size(C.c_uint(int(bits)))

    def run_chain(self, stages, data, nbits=None, vs=32, nf=100.0, adaptive=1, out_cap=None):
        src = _as_u8(data)
        if nbits is None:
            nbits = src.size * 8
        if out_cap is None:
            out_cap = max(4096, src.size * 24 + 4096)
        names = [STAGE_TO_CLI[s][0].encode() for s in stages]
        enc = [STAGE_TO_CLI[s][1] for s in stages]
        c_names = (C.c_char_p * len(stages))(*names)
        c_enc = (C.c_int * len(stages))(*enc)
        out = np.zeros(out_cap, dtype=np.uint8)
        out_bits = C.c_ulonglong(0)
        failed = C.c_int(-1)
        rc = self.lib.ref_run_chain(c_names, c_enc, len(stages), src.ctypes.data if src.size else None, nbits, vs, nf,
                                    int(adaptive), out.ctypes.data, out_cap, C.byref(out_bits), C.byref(failed))
        nb = int(out_bits.value)
        return int(rc), out[: (nb + 7) // 8].tobytes(), nb

    def dega_encode(self, samples, vs=32, nf=100.0, adaptive=1):
        x = np.ascontiguousarray(samples, dtype=np.float32)
        return self.run_chain(DEGA_ENC, x, x.size * 32, vs, nf, adaptive)

    def dega_decode(self, payload, nbits=None, vs=32, nf=100.0, adaptive=1, out_cap=None):
        rc, raw, nb = self.run_chain(DEGA_DEC, payload, nbits, vs, nf, adaptive, out_cap=out_cap)
        return rc, np.frombuffer(raw[: (nb // 32) * 4], dtype=np.float32).copy()
