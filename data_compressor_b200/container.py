"""On-disk container of a batch encode: header + bit-length index + the streams back to back.

Every stream's byte range is byte-identical to the file the reference's DCCLI writes for the same samples and
chain (DCIOLib/src/bit_file_buffer.c:310-333), so one range extracted with `dd` decodes with the reference's own
`decode bac adaptive # decode seg # decode diff # decode normalize`.  The file format and the position-based writers
are C (csrc/container.cpp, include/dc_cuda.h); this module only moves tensors in and out.  No GPU is needed to read
or write a container; decoding one is Context.dega_decode's job.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import ContainerHeader, DcError, Layout, Params

CHAIN_DEGA, CHAIN_AXDR, CHAIN_LZMH = 1, 2, 3


def _check(rc: int, what: str):
    if rc != 0:
        raise DcError(rc, what)


def make_header(chain: int, n_streams: int, payload_bytes: int, n_channels: int = 0, channel_len: int = 0, block_len: int = 0,
                vs: int = 32, nf: float = 100.0, adaptive: bool = True) -> ContainerHeader:
    lib = _lib.load()
    h = ContainerHeader()
    par = Params(int(vs), float(nf), int(bool(adaptive)), 0)
    lay = Layout(int(n_channels), int(channel_len), int(block_len), 0, None)
    _check(lib.dc_container_init_header(C.byref(h), int(chain), C.byref(par), C.byref(lay), int(n_streams), int(payload_bytes)),
           "dc_container_init_header")
    return h


def create(path: str, header: ContainerHeader) -> None:
    """Creates the file with its header and final size (one rank of a job does this, before any write_shard)."""
    _check(_lib.load().dc_container_create(os.fsencode(path), C.byref(header)), "dc_container_create")


def write_shard(path: str, header: ContainerHeader, first_stream: int, bit_len, byte_base: int, payload) -> None:
    """Writes one shard's index entries and payload bytes at their container positions (host arrays)."""
    bl = np.ascontiguousarray(bit_len, dtype=np.uint64)
    pay = np.ascontiguousarray(payload, dtype=np.uint8)
    _check(_lib.load().dc_container_write_shard(os.fsencode(path), C.byref(header), int(first_stream), bl.size,
                                                bl.ctypes.data if bl.size else None, int(byte_base),
                                                pay.ctypes.data if pay.size else None, pay.size), "dc_container_write_shard")


def write_encoded(path: str, enc, chain: int = CHAIN_DEGA, vs: int = 32, nf: float = 100.0, adaptive: bool = True) -> ContainerHeader:
    """Single process: writes a codec.Encoded (device tensors) into `path`."""
    h, bit_len, payload = prepare_encoded(path, enc, chain, vs, nf, adaptive, None)
    write_encoded_shard(path, h, None, bit_len, payload)
    return h


def prepare_encoded(path: str, enc, chain: int = CHAIN_DEGA, vs: int = 32, nf: float = 100.0, adaptive: bool = True,
                    sizes: dict | None = None):
    """First half of a (possibly multi-rank) write: builds the job-wide header, rank 0 creates the file.

    `sizes` is the dict shard.exchange_sizes returned for this rank (None = single process); `enc.n_channels` is this
    rank's channel count, the header carries the job's totals.  The caller puts a barrier between this and
    write_encoded_shard.  Returns (header, this rank's bit_len, this rank's payload) as host arrays."""
    n_local = enc.n_streams
    bit_len = enc.bit_len.cpu().numpy().astype(np.uint64)
    payload = enc.payload[: enc.total_bytes].cpu().numpy()
    if sizes is None:
        sizes = {"rank": 0, "world": 1, "byte_base": 0, "stream_base": 0, "total_bytes": enc.total_bytes, "total_streams": n_local,
                 "counts": [n_local]}
    blocks_per_channel = max(1, n_local // max(1, enc.n_channels))
    total_channels = sizes["total_streams"] // blocks_per_channel
    h = make_header(chain, sizes["total_streams"], sizes["total_bytes"], total_channels, enc.channel_len, enc.block_len, vs, nf, adaptive)
    if sizes["rank"] == 0:
        create(path, h)
    return h, bit_len, payload


def write_encoded_shard(path: str, header: ContainerHeader, sizes: dict | None, bit_len, payload) -> None:
    """Second half of a multi-rank write: after the barrier that follows rank 0's create()."""
    sizes = sizes or {"stream_base": 0, "byte_base": 0}
    write_shard(path, header, sizes["stream_base"], bit_len, sizes["byte_base"], payload)


def read_header(path: str) -> ContainerHeader:
    h = ContainerHeader()
    _check(_lib.load().dc_container_read_header(os.fsencode(path), C.byref(h)), "dc_container_read_header")
    return h


def read_index(path: str, header: ContainerHeader | None = None, first_stream: int = 0, n: int | None = None) -> np.ndarray:
    h = header or read_header(path)
    n = h.n_streams - first_stream if n is None else n
    out = np.empty(n, dtype=np.uint64)
    _check(_lib.load().dc_container_read_index(os.fsencode(path), C.byref(h), int(first_stream), int(n),
                                               out.ctypes.data if n else None), "dc_container_read_index")
    return out


def read_payload(path: str, header: ContainerHeader | None = None, byte_off: int = 0, nbytes: int | None = None) -> np.ndarray:
    h = header or read_header(path)
    nbytes = h.payload_bytes - byte_off if nbytes is None else nbytes
    out = np.empty(nbytes, dtype=np.uint8)
    _check(_lib.load().dc_container_read_payload(os.fsencode(path), C.byref(h), int(byte_off), int(nbytes),
                                                 out.ctypes.data if nbytes else None), "dc_container_read_payload")
    return out


def byte_offsets(bit_len: np.ndarray) -> np.ndarray:
    """Exclusive prefix sum of ceil(bit_len / 8): where every stream starts inside the payload (n + 1 entries)."""
    nbytes = (np.asarray(bit_len, dtype=np.uint64) + np.uint64(7)) >> np.uint64(3)
    off = np.zeros(nbytes.size + 1, dtype=np.uint64)
    np.cumsum(nbytes, out=off[1:])
    return off


def stream_file(path: str, s: int) -> tuple[bytes, int]:
    """The bytes DCCLI would have written for stream s (what `dd` with these offsets extracts) and its bit count."""
    h = read_header(path)
    bl = read_index(path, h, 0, s + 1)
    off = int(byte_offsets(bl[:s])[-1]) if s else 0
    nb = int(bl[s])
    return read_payload(path, h, off, (nb + 7) // 8).tobytes(), nb


def load_encoded(path: str, device="cuda"):
    """Reads a whole container into a codec.Encoded on `device` (plus its header), ready for Context.dega_decode."""
    import torch
    from .codec import Encoded
    h = read_header(path)
    bl = read_index(path, h)
    off = byte_offsets(bl)
    if int(off[-1]) != int(h.payload_bytes):   # a truncated or corrupt index: the decoders would read past the payload
        raise DcError(-3, f"container index covers {int(off[-1])} payload bytes, header says {int(h.payload_bytes)}")
    pay = read_payload(path, h)
    payload = torch.empty(pay.size + 16, dtype=torch.uint8, device=device)
    payload[: pay.size] = torch.from_numpy(pay).to(device)
    enc = Encoded(payload=payload, bit_len=torch.from_numpy(bl.astype(np.int64)).to(device),
                  byte_off=torch.from_numpy(off.astype(np.int64)).to(device),
                  status=torch.zeros(bl.size, dtype=torch.int32, device=device), total_bytes=int(h.payload_bytes),
                  n_channels=int(h.n_channels), channel_len=int(h.channel_len), block_len=int(h.block_len))
    return enc, h
