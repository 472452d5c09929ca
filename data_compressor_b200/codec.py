"""Batch face of the hot path: many independent streams per call, device-resident (torch tensors).

torch is plumbing here (device memory, streams, events); every codec result comes from the hand-written sm_100a
kernels behind include/dc_cuda.h.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import DcError, Layout, Params


@dataclass
class Encoded:
    """Container-ready result of a batch encode (all tensors on the GPU).

    payload[byte_off[s] : byte_off[s] + ceil(bit_len[s] / 8)] is byte-identical to the file DCCLI writes for
    stream s (DCIOLib/src/bit_file_buffer.c:310-333)."""
    payload: torch.Tensor      # uint8 [capacity]; first total_bytes are valid
    bit_len: torch.Tensor      # int64 [n_streams]
    byte_off: torch.Tensor     # int64 [n_streams + 1]
    status: torch.Tensor       # int32 [n_streams]
    total_bytes: int
    n_channels: int
    channel_len: int
    block_len: int

    @property
    def n_streams(self) -> int:
        return int(self.bit_len.numel())

    def stream_bytes(self, s: int) -> bytes:
        o = int(self.byte_off[s])
        n = (int(self.bit_len[s]) + 7) // 8
        return bytes(self.payload[o:o + n].cpu().numpy().tobytes())


class Pending:
    """A *_host_async job queued on a context.  Jobs are collected in submission order (dc_wait_next): wait() first
    collects the jobs queued before this one, then this one, and returns what the blocking call returns."""

    def __init__(self, ctx, what, result, keep=()):
        self.ctx, self.what, self.result, self.keep = ctx, what, result, keep   # keep: the buffers the job reads
        self.rc = None
        ctx.__dict__.setdefault("_pending", []).append(self)

    def wait(self):
        q = self.ctx._pending
        while self.rc is None:
            head = q.pop(0)
            head.rc = int(self.ctx.lib.dc_wait_next(self.ctx.h))
            head.keep = ()
        self.ctx._check(self.rc, self.what)
        return self.result()


class Context:
    """One dc_ctx: one GPU, one CUDA stream, one scratch arena.  Not re-entrant (like the reference codecs)."""

    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise DcError(_lib.ERR_LIBRARY_INIT, "no CUDA device: data_compressor_b200 has no CPU fallback")
        self.lib = _lib.load()
        self.device = int(device)
        h = C.c_void_p()
        rc = self.lib.dc_create(C.byref(h), self.device)
        if rc != 0:
            raise DcError(rc, "dc_create")
        self.h = h
        self.torch_device = torch.device("cuda", self.device)
        self.stream = torch.cuda.ExternalStream(self.lib.dc_stream(self.h), device=self.torch_device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.dc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise DcError(rc, f"{what}: {self.lib.dc_last_error(self.h).decode()}")

    def launch_count(self) -> int:
        return int(self.lib.dc_launch_count(self.h))

    def coder_wave_streams(self) -> int:
        """Streams the coder kernels hold in flight at once on this GPU (dc_coder_wave_streams)."""
        return int(self.lib.dc_coder_wave_streams(self.h))

    def two_wave_block_len(self, n_channels: int, channel_len: int) -> int:
        """Shortest block length whose stream count fits two waves of the coder kernels (whole channels if even those
        do not fill them)."""
        per_channel = max(1, 2 * self.coder_wave_streams() // max(1, n_channels))
        return 0 if per_channel <= 1 else -(-channel_len // per_channel)

    def set_profiling(self, enable: bool):
        self._check(self.lib.dc_set_profiling(self.h, 1 if enable else 0), "dc_set_profiling")

    def kernel_ms(self) -> dict:
        """Device time (ms) of each kernel group during the last call (CUDA events on the context's stream)."""
        arr = (C.c_float * len(_lib.KERNEL_NAMES))()
        self._check(self.lib.dc_get_kernel_ms(self.h, arr), "dc_get_kernel_ms")
        return {k: float(arr[i]) for i, k in enumerate(_lib.KERNEL_NAMES)}

    def synchronize(self):
        self._check(self.lib.dc_synchronize(self.h), "dc_synchronize")

    @staticmethod
    def _params(vs: int, nf: float, adaptive: bool, blocksize: int = 0) -> Params:
        return Params(int(vs), float(nf), 1 if adaptive else 0, int(blocksize))

    @staticmethod
    def _layout(n_channels: int, channel_len: int, block_len: int) -> Layout:
        return Layout(int(n_channels), int(channel_len), int(block_len), 0, None)

    def n_streams(self, n_channels: int, channel_len: int, block_len: int) -> int:
        lay = self._layout(n_channels, channel_len, block_len)
        return int(self.lib.dc_layout_streams(C.byref(lay)))

    # ------------------------------------------------------------------------------------------------
    def _encode(self, fn, what, samples, block_len, vs, nf, adaptive, payload, payload_cap, out):
        assert samples.is_cuda and samples.dtype == torch.float32 and samples.is_contiguous() and samples.dim() == 2
        n_channels, channel_len = samples.shape
        n = self.n_streams(n_channels, channel_len, block_len)
        dev = samples.device
        if out is None:
            if payload_cap is None:
                # generous default: raw size + per-stream slack (DEGA never expands REDD-like data)
                payload_cap = samples.numel() * 4 + n * 16 + 64
            payload = torch.empty(int(payload_cap) + 16, dtype=torch.uint8, device=dev)
            bit_len = torch.empty(n, dtype=torch.int64, device=dev)
            byte_off = torch.empty(n + 1, dtype=torch.int64, device=dev)
            status = torch.empty(n, dtype=torch.int32, device=dev)
        else:
            payload, bit_len, byte_off, status = out.payload, out.bit_len, out.byte_off, out.status
            payload_cap = payload.numel() - 16
        lay = self._layout(n_channels, channel_len, block_len)
        par = self._params(vs, nf, adaptive)
        total = C.c_uint64(0)
        torch.cuda.current_stream(dev).synchronize()
        rc = fn(self.h, samples.data_ptr(), C.byref(lay), C.byref(par), payload.data_ptr(), int(payload_cap),
                bit_len.data_ptr(), byte_off.data_ptr(), status.data_ptr(), C.byref(total))
        if rc == _lib.ERR_MEMORY and total.value > payload_cap:
            # the library reports the size it needed (series that DEGA expands, e.g. coarse granularities with large
            # deltas): allocate exactly that and run once more
            payload_cap = int(total.value)
            payload = torch.empty(payload_cap + 16, dtype=torch.uint8, device=dev)
            rc = fn(self.h, samples.data_ptr(), C.byref(lay), C.byref(par), payload.data_ptr(), payload_cap,
                    bit_len.data_ptr(), byte_off.data_ptr(), status.data_ptr(), C.byref(total))
        self._check(rc, what)
        return Encoded(payload, bit_len, byte_off, status, int(total.value), n_channels, channel_len, int(block_len))

    def dega_encode(self, samples: torch.Tensor, block_len: int = 0, vs: int = 32, nf: float = 100.0,
                    adaptive: bool = True, payload_cap: int | None = None, out: Encoded | None = None) -> Encoded:
        """`encode normalize # encode diff # encode seg # encode bac [adaptive]` for every (channel, block).

        samples: float32 [n_channels, channel_len] on the GPU (one raw float32 "file" per channel);
        block_len: samples per resumable block, 0 = whole channel."""
        return self._encode(self.lib.dc_dega_encode_device, "dc_dega_encode_device", samples, block_len, vs, nf,
                            adaptive, None, payload_cap, out)

    def axdr_encode(self, samples: torch.Tensor, block_len: int = 0, vs: int = 32, nf: float = 100.0,
                    payload_cap: int | None = None, out: Encoded | None = None) -> Encoded:
        """`encode normalize` (A-XDR) for every stream."""
        return self._encode(self.lib.dc_axdr_encode_device, "dc_axdr_encode_device", samples, block_len, vs, nf,
                            False, None, payload_cap, out)

    def _decode(self, fn, what, enc: Encoded, vs, nf, adaptive, out):
        dev = enc.payload.device
        n = enc.n_streams
        if out is None:
            out = torch.empty((enc.n_channels, enc.channel_len), dtype=torch.float32, device=dev)
        n_out = torch.empty(n, dtype=torch.int64, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        lay = self._layout(enc.n_channels, enc.channel_len, enc.block_len)
        par = self._params(vs, nf, adaptive)
        torch.cuda.current_stream(dev).synchronize()
        rc = fn(self.h, enc.payload.data_ptr(), int(enc.total_bytes), enc.bit_len.data_ptr(), enc.byte_off.data_ptr(),
                C.byref(lay), C.byref(par), out.data_ptr(), n_out.data_ptr(), status.data_ptr())
        self._check(rc, what)
        self.synchronize()
        return out, n_out, status

    def dega_decode(self, enc: Encoded, vs: int = 32, nf: float = 100.0, adaptive: bool = True,
                    out: torch.Tensor | None = None):
        """`decode bac [adaptive] # decode seg # decode diff # decode normalize` for every stream.

        Returns (samples [n_channels, channel_len], n_out [n_streams], status [n_streams])."""
        return self._decode(self.lib.dc_dega_decode_device, "dc_dega_decode_device", enc, vs, nf, adaptive, out)

    def axdr_decode(self, enc: Encoded, vs: int = 32, nf: float = 100.0, out: torch.Tensor | None = None):
        return self._decode(self.lib.dc_axdr_decode_device, "dc_axdr_decode_device", enc, vs, nf, False, out)

    # ---- csv: text column -> float32 --------------------------------------------------------------
    def csv_decode(self, text, column: int = 1, separator: str = ",") -> np.ndarray:
        """`decode csv column=.. separator=..` (DCLib/src/csv.c:13-44): the column's values as float32, in line order."""
        src = np.frombuffer(text.encode() if isinstance(text, str) else bytes(text), dtype=np.uint8)
        cap = max(16, int(np.count_nonzero(src == 10)) + 1)
        out = np.empty(cap, dtype=np.float32)
        n = C.c_uint64(0)
        rc = self.lib.dc_csv_decode_host(self.h, src.ctypes.data if src.size else None, src.size, int(column), separator.encode()[:1],
                                         out.ctypes.data, cap, C.byref(n))
        self._check(rc, "dc_csv_decode_host")
        return out[: n.value].copy()

    def csv_encode(self, values, column: int = 1, separator: str = ",", decimals: int = 2) -> bytes:
        """`encode csv column=.. separator_char=.. num_decimal_places=..` (DCLib/src/csv.c:46-65): the text of the values."""
        x = np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
        cap = x.size * (column + 15) + 64               # values below 10^7 fit; larger ones take the retry
        n = C.c_uint64(0)
        for _ in range(2):
            out = np.empty(cap, dtype=np.uint8)
            rc = self.lib.dc_csv_encode_host(self.h, x.ctypes.data if x.size else None, x.size, int(column), separator.encode()[:1], int(decimals),
                                             out.ctypes.data, cap, C.byref(n))
            if rc != _lib.ERR_MEMORY or n.value <= cap:
                break
            cap = int(n.value)
        self._check(rc, "dc_csv_encode_host")
        return out[: n.value].tobytes()

    # ---- aggregate: granularity reduction ---------------------------------------------------------
    def aggregate(self, x: torch.Tensor, num_values: int) -> torch.Tensor:
        """`encode aggregate num_values=k` (DCLib/src/aggregate.c:9-27) for every channel of a float32
        [n_channels, channel_len] tensor: sums of k consecutive samples, added left to right; the last group of a
        channel may be short.  Returns [n_channels, ceil(channel_len / k)]."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2 and num_values >= 1
        nc, cl = x.shape
        out = torch.empty((nc, (cl + num_values - 1) // num_values), dtype=torch.float32, device=x.device)
        torch.cuda.current_stream(x.device).synchronize()
        rc = self.lib.dc_aggregate_device(self.h, x.data_ptr(), nc, cl, int(num_values), out.data_ptr())
        self._check(rc, "dc_aggregate_device")
        self.synchronize()
        return out

    # ---- LZMH: independent byte blocks ------------------------------------------------------------
    def lzmh_encode(self, data: torch.Tensor, block_bytes: int = 65536, payload_cap: int | None = None):
        """`encode lzmh` for every block of `block_bytes` bytes of a uint8 tensor (last block may be short).
        Returns (payload, bit_len, byte_off, status, total_bytes, in_off)."""
        assert data.is_cuda and data.dtype == torch.uint8 and data.is_contiguous()
        dev = data.device
        nbytes = data.numel()
        n = (nbytes + block_bytes - 1) // block_bytes if nbytes else 0
        in_off = torch.clamp(torch.arange(n + 1, dtype=torch.int64, device=dev) * block_bytes, max=nbytes)
        if payload_cap is None:
            payload_cap = nbytes * 10 // 8 + n * 16 + 64
        payload = torch.empty(int(payload_cap) + 16, dtype=torch.uint8, device=dev)
        bit_len = torch.empty(n, dtype=torch.int64, device=dev)
        byte_off = torch.empty(n + 1, dtype=torch.int64, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        total = C.c_uint64(0)
        torch.cuda.current_stream(dev).synchronize()
        rc = self.lib.dc_lzmh_encode_device(self.h, data.data_ptr(), in_off.data_ptr(), n, payload.data_ptr(), int(payload_cap),
                                            bit_len.data_ptr(), byte_off.data_ptr(), status.data_ptr(), C.byref(total))
        self._check(rc, "dc_lzmh_encode_device")
        return payload, bit_len, byte_off, status, int(total.value), in_off

    def lzmh_decode(self, payload: torch.Tensor, total_bytes: int, bit_len: torch.Tensor, byte_off: torch.Tensor,
                    out_off: torch.Tensor):
        """`decode lzmh` for every block; out_off[n+1] gives each block's output window."""
        dev = payload.device
        n = int(bit_len.numel())
        out = torch.empty(int(out_off[-1].item()) if n else 0, dtype=torch.uint8, device=dev)
        out_len = torch.empty(n, dtype=torch.int64, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        torch.cuda.current_stream(dev).synchronize()
        rc = self.lib.dc_lzmh_decode_device(self.h, payload.data_ptr(), int(total_bytes), bit_len.data_ptr(), byte_off.data_ptr(), n,
                                            out.data_ptr(), out_off.data_ptr(), out_len.data_ptr(), status.data_ptr())
        self._check(rc, "dc_lzmh_decode_device")
        self.synchronize()
        return out, out_len, status

    # ---- host-buffer face (numpy in / numpy out; copies inside) -----------------------------------
    def _pinned(self, key: str, nbytes: int) -> np.ndarray:
        """Reusable page-locked host buffer (uint8 view); the *_host calls copy with cudaMemcpyAsync."""
        cache = self.__dict__.setdefault("_pinned_cache", {})
        buf = cache.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, pin_memory=True)
            cache[key] = buf
        return buf.numpy()[:nbytes]

    def dega_encode_host(self, samples: np.ndarray, block_len: int = 0, vs: int = 32, nf: float = 100.0,
                         adaptive: bool = True, payload_cap: int | None = None, wait: bool = True, slot: int = 0):
        """Host-buffer face (dc_dega_encode_host).  Outputs live in reusable pinned buffers owned by the context:
        they stay valid until the next *_host encode call.

        wait=False starts the call (dc_dega_encode_host_async) and returns a Pending whose wait() gives the same
        tuple; until then neither `samples` nor the context may be touched.  An encode in flight on one context next
        to a decode on another uses the PCIe link in both directions at once.  `slot` picks one of several sets of
        output buffers, so that the result of one call can still be read (e.g. decoded) while the next call runs."""
        x = np.ascontiguousarray(samples, dtype=np.float32)
        assert x.ndim == 2
        n_channels, channel_len = x.shape
        n = self.n_streams(n_channels, channel_len, block_len)
        if payload_cap is None:
            payload_cap = x.size * 4 + n * 16 + 64
        payload = self._pinned(f"enc_payload{slot}", payload_cap)
        bit_len = self._pinned(f"enc_bit_len{slot}", n * 8).view(np.uint64)
        byte_off = self._pinned(f"enc_byte_off{slot}", (n + 1) * 8).view(np.uint64)
        status = self._pinned(f"enc_status{slot}", n * 4).view(np.int32)
        total = C.c_uint64(0)
        lay = self._layout(n_channels, channel_len, block_len)
        par = self._params(vs, nf, adaptive)
        fn = self.lib.dc_dega_encode_host if wait else self.lib.dc_dega_encode_host_async
        rc = fn(self.h, x.ctypes.data, C.byref(lay), C.byref(par), payload.ctypes.data, payload_cap, bit_len.ctypes.data,
                byte_off.ctypes.data, status.ctypes.data, C.byref(total))
        self._check(rc, "dc_dega_encode_host")
        result = lambda: (payload[: total.value], bit_len, byte_off, status)
        return result() if wait else Pending(self, "dc_dega_encode_host_async", result, keep=(x, total))

    def dega_decode_host(self, payload: np.ndarray, bit_len: np.ndarray, byte_off: np.ndarray, n_channels: int,
                         channel_len: int, block_len: int = 0, vs: int = 32, nf: float = 100.0, adaptive: bool = True,
                         wait: bool = True):
        """Host-buffer face (dc_dega_decode_host); wait=False: see dega_encode_host."""
        p = np.ascontiguousarray(payload, dtype=np.uint8)
        bl = np.ascontiguousarray(bit_len, dtype=np.uint64)
        bo = np.ascontiguousarray(byte_off, dtype=np.uint64)
        n = bl.size
        out = self._pinned("dec_samples", n_channels * channel_len * 4).view(np.float32).reshape(n_channels, channel_len)
        n_out = self._pinned("dec_n_out", n * 8).view(np.uint64)
        status = self._pinned("dec_status", n * 4).view(np.int32)
        lay = self._layout(n_channels, channel_len, block_len)
        par = self._params(vs, nf, adaptive)
        fn = self.lib.dc_dega_decode_host if wait else self.lib.dc_dega_decode_host_async
        rc = fn(self.h, p.ctypes.data, p.size, bl.ctypes.data, bo.ctypes.data, C.byref(lay), C.byref(par), out.ctypes.data,
                n_out.ctypes.data, status.ctypes.data)
        self._check(rc, "dc_dega_decode_host")
        return (out, n_out, status) if wait else Pending(self, "dc_dega_decode_host_async", lambda: (out, n_out, status), keep=(p, bl, bo))

    # ---- plugin face: one stage, one stream, host bytes -------------------------------------------
    def stage_run(self, stage: int, data, nbits: int | None = None, vs: int = 32, nf: float = 100.0,
                  adaptive: bool = False, blocksize: int = 0):
        """Runs ONE reference stage on the GPU (what an enc_dec_function_t body calls).  Returns
        (code, bytes, nbits): code is 0 or the negative err_codes.h value the reference stage returns."""
        src = np.frombuffer(bytes(data), dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) \
            else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        if nbits is None:
            nbits = src.size * 8
        par = self._params(vs, nf, adaptive, blocksize)
        cap = int(self.lib.dc_stage_out_bound(stage, nbits, C.byref(par)))
        out_bits = C.c_uint64(0)
        for _ in range(3):
            out = np.empty(max(cap, 16), dtype=np.uint8)
            rc = self.lib.dc_stage_run_host(self.h, stage, src.ctypes.data if src.size else None, nbits, C.byref(par),
                                            out.ctypes.data, out.size, C.byref(out_bits))
            if rc == _lib.ERR_MEMORY and out_bits.value and (out_bits.value + 7) // 8 > out.size:
                cap = (out_bits.value + 7) // 8 + 16     # exact size reported: retry once
                continue
            break
        nb = int(out_bits.value) if rc == 0 else 0
        return rc, out[: (nb + 7) // 8].tobytes(), nb

    def chain_run(self, stages, data, nbits: int | None = None, vs: int = 32, nf: float = 100.0, adaptive: bool = False,
                  out_cap: int | None = None, blocksize: int = 0):
        """`encode X # encode Y # ...` over one stream with the intermediate results kept on the device
        (dc_chain_run_host).  `stages` is a list of stage numbers; the options apply to every stage, as when the
        command line repeats them.  Returns (code, bytes, nbits, failed_stage)."""
        src = np.frombuffer(bytes(data), dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) \
            else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        if nbits is None:
            nbits = src.size * 8
        n = len(stages)
        st = (C.c_int * n)(*[int(x) for x in stages])
        pars = (Params * n)(*[self._params(vs, nf, adaptive, blocksize) for _ in range(n)])
        cap = out_cap or max(1 << 16, src.size * 40 + 4096)
        out_bits, failed = C.c_uint64(0), C.c_uint32(0)
        for _ in range(3):
            out = np.empty(cap, dtype=np.uint8)
            rc = self.lib.dc_chain_run_host(self.h, n, st, pars, src.ctypes.data if src.size else None, nbits, out.ctypes.data, out.size,
                                            C.byref(out_bits), C.byref(failed))
            if rc == _lib.ERR_MEMORY and out_bits.value and (out_bits.value + 7) // 8 > out.size:
                cap = (out_bits.value + 7) // 8 + 16
                continue
            break
        nb = int(out_bits.value) if rc == 0 else 0
        return rc, out[: (nb + 7) // 8].tobytes(), nb, int(failed.value)
