"""ctypes binding of libdc_cuda.so (C ABI: include/dc_cuda.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C data_compressor_b200/csrc``.  There is no
Python or CPU implementation behind these calls: if the shared object is missing, or no GPU is usable, the
package raises instead of falling back.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libdc_cuda.so")
CSRC = os.path.join(HERE, "csrc")

# common/inc/err_codes.h:8-15,28-29
OK = 0
ERR_INVALID_VALUE = -1
ERR_VALUE_TOO_LARGE = -2
ERR_INVALID_FORMAT = -3
ERR_INVALID_MODE = -4
ERR_FILE_IO = -5
ERR_MEMORY = -6
ERR_LIBRARY_INIT = -10
ERR_LIBRARY_CALL = -11

ERROR_MESSAGES = {
    0: "Successful", -1: "Invalid value", -2: "Value larger than usable size", -3: "Invalid format",
    -4: "Invalid mode", -5: "File I/O error", -6: "Memory error",
    -10: "Error initializing library", -11: "Error calling library",
}

# enum dc_stage (include/dc_cuda.h)
(STAGE_NORMALIZE, STAGE_DENORMALIZE, STAGE_DIFF_ENC, STAGE_DIFF_DEC, STAGE_SEG_ENC, STAGE_SEG_DEC,
 STAGE_BAC_ENC, STAGE_BAC_DEC, STAGE_LZMH_ENC, STAGE_LZMH_DEC, STAGE_DEGA_ENC, STAGE_DEGA_DEC, STAGE_COPY) = range(13)


KERNEL_NAMES = ("front", "bac_encode", "scan", "gather", "bac_decode", "back", "lzmh_encode", "lzmh_decode")


class Params(C.Structure):
    """dc_params: the subset of options_t (DCLib/inc/enc_dec.h:28-41) the hot path reads."""
    _fields_ = [("value_size_bits", C.c_uint32), ("normalization_factor", C.c_float),
                ("adaptive", C.c_int32), ("block_size_bits", C.c_uint32)]


class ContainerHeader(C.Structure):
    """dc_container_header (include/dc_cuda.h): 128 bytes, little endian."""
    _fields_ = [("magic", C.c_char * 8), ("version", C.c_uint32), ("chain", C.c_uint32), ("value_size_bits", C.c_uint32),
                ("normalization_factor", C.c_float), ("adaptive", C.c_int32), ("reserved0", C.c_uint32),
                ("n_channels", C.c_uint64), ("channel_len", C.c_uint64), ("block_len", C.c_uint64), ("n_streams", C.c_uint64),
                ("payload_bytes", C.c_uint64), ("index_offset", C.c_uint64), ("payload_offset", C.c_uint64),
                ("reserved", C.c_uint64 * 5)]


class Layout(C.Structure):
    """dc_layout: where the streams live inside one flat, channel-major sample array."""
    _fields_ = [("n_channels", C.c_uint64), ("channel_len", C.c_uint64), ("block_len", C.c_uint64),
                ("n_streams", C.c_uint64), ("sample_off", C.c_void_p)]


#: every symbol include/dc_cuda.h declares: name -> (restype, argtypes)
_P, _U64, _I = C.c_void_p, C.c_uint64, C.c_int
_PL, _PP = C.POINTER(Layout), C.POINTER(Params)
_PU64 = C.POINTER(C.c_uint64)
_PH = C.POINTER(ContainerHeader)
SYMBOLS = {
    "dc_abi_version": (_I, []),
    "dc_create": (_I, [C.POINTER(C.c_void_p), _I]),
    "dc_destroy": (None, [_P]),
    "dc_last_error": (C.c_char_p, [_P]),
    "dc_layout_streams": (_U64, [_PL]),
    "dc_synchronize": (_I, [_P]),
    "dc_stream": (_P, [_P]),
    "dc_launch_count": (_U64, [_P]),
    "dc_coder_wave_streams": (_U64, [_P]),
    "dc_set_profiling": (_I, [_P, _I]),
    "dc_get_kernel_ms": (_I, [_P, C.POINTER(C.c_float)]),
    "dc_dega_encode_device": (_I, [_P, _P, _PL, _PP, _P, _U64, _P, _P, _P, _PU64]),
    "dc_dega_decode_device": (_I, [_P, _P, _U64, _P, _P, _PL, _PP, _P, _P, _P]),
    "dc_dega_encode_host": (_I, [_P, _P, _PL, _PP, _P, _U64, _P, _P, _P, _PU64]),
    "dc_dega_decode_host": (_I, [_P, _P, _U64, _P, _P, _PL, _PP, _P, _P, _P]),
    "dc_dega_encode_host_async": (_I, [_P, _P, _PL, _PP, _P, _U64, _P, _P, _P, _PU64]),
    "dc_dega_decode_host_async": (_I, [_P, _P, _U64, _P, _P, _PL, _PP, _P, _P, _P]),
    "dc_wait": (_I, [_P]),
    "dc_wait_next": (_I, [_P]),
    "dc_axdr_encode_device": (_I, [_P, _P, _PL, _PP, _P, _U64, _P, _P, _P, _PU64]),
    "dc_axdr_decode_device": (_I, [_P, _P, _U64, _P, _P, _PL, _PP, _P, _P, _P]),
    "dc_lzmh_encode_device": (_I, [_P, _P, _P, _U64, _P, _U64, _P, _P, _P, _PU64]),
    "dc_lzmh_decode_device": (_I, [_P, _P, _U64, _P, _P, _U64, _P, _P, _P, _P]),
    "dc_aggregate_device": (_I, [_P, _P, _U64, _U64, C.c_uint32, _P]),
    "dc_csv_decode_host": (_I, [_P, _P, _U64, C.c_uint32, C.c_char, _P, _U64, _PU64]),
    "dc_csv_encode_host": (_I, [_P, _P, _U64, C.c_uint32, C.c_char, C.c_uint32, _P, _U64, _PU64]),
    "dc_container_init_header": (_I, [_PH, C.c_uint32, _PP, _PL, _U64, _U64]),
    "dc_container_create": (_I, [C.c_char_p, _PH]),
    "dc_container_write_shard": (_I, [C.c_char_p, _PH, _U64, _U64, _P, _U64, _P, _U64]),
    "dc_container_read_header": (_I, [C.c_char_p, _PH]),
    "dc_container_read_index": (_I, [C.c_char_p, _PH, _U64, _U64, _P]),
    "dc_container_read_payload": (_I, [C.c_char_p, _PH, _U64, _U64, _P]),
    "dc_stage_run_host": (_I, [_P, _I, _P, _U64, _PP, _P, _U64, _PU64]),
    "dc_chain_run_host": (_I, [_P, C.c_uint32, C.POINTER(_I), _PP, _P, _U64, _P, _U64, _PU64, C.POINTER(C.c_uint32)]),
    "dc_stage_out_bound": (_U64, [_I, _U64, _PP]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libdc_cuda.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", CSRC] + ([] if verbose else ["-s"]), check=True)
    return LIB_PATH


def load() -> C.CDLL:
    """Load the shared library and bind every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` or "
            f"`make -C {CSRC}`.  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError here == ABI mismatch with include/dc_cuda.h
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DcError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        msg = ERROR_MESSAGES.get(code, "Unknown error")
        super().__init__(f"{msg} ({code})" + (f": {detail}" if detail else ""))
