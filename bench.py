#!/usr/bin/env python
"""bench.py -- DEGA encode+decode throughput on B200 (BASELINE.json metric) with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W                 # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU path on the host cores

Workload = BASELINE.json configs[1]: synthetic REDD-like 1 Hz load profiles, 4096 channels x 1 day per GPU,
DEGA (normalize -> diff -> seg -> bac adaptive) encode + decode, channels cut into resumable blocks.  The block
length is derived from the device (Context.two_wave_block_len: the shortest block whose stream count fits two waves
of the coder kernels, 1168 samples on a 148-SM B200); the line also carries the same step at 1024-sample blocks and
with whole channels (the reference's own unit of work), so the fitted figure is not the only one.
A "step" is one encode + one decode of every stream on the rank.  `value` counts the raw float32 bytes once per
step (GB/s of raw samples through the encode+decode round trip), inputs resident in HBM; `e2e` is the same
step through the host-buffer C ABI (dc_dega_encode_host[_async] / dc_dega_decode_host[_async]) with pinned host memory.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHANNELS = 4096
SAMPLES = 86400
BLOCK_LEN = 1168   # what Context.two_wave_block_len gives on a 148-SM B200 (74 blocks per channel-day -> 303104 streams = 2 waves of
                   # 148 SMs x 1024 threads); only used where no GPU context exists to ask (the --impl reference arm off a GPU box)
VS, NF, ADAPTIVE = 32, 100.0, True
METRIC = "DEGA encode+decode GB/s of raw samples"


def workload_name(channels, samples, block_len):
    return (f"configs[1]: synthetic REDD-like 1 Hz load profiles, {channels} channels x {samples} samples per GPU, DEGA "
            f"(normalize->diff->seg->bac adaptive) encode+decode, block_len {block_len}, valuesize {VS}, normalization_factor {NF:g}")


# --------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons DURING the timed region through NVML (in-process: the nvidia-smi CLI in
    loop mode stalls the driver for milliseconds per sample, which showed up as 5-10 ms of extra step time)."""

    def __init__(self, gpu_index, period=0.1):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.period = period
        self.rows = []
        self.stop_flag = False
        self.max_mhz = None
        self.ok = False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.gpu]) if visible and visible.split(",")[self.gpu].isdigit() else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.ok = True
            while not self.stop_flag:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((sm, reasons))
                time.sleep(self.period)
        except Exception as ex:   # no NVML: report it, never fail the bench
            self.rows.append((None, repr(ex)))

    def stop(self):
        self.stop_flag = True

    def summary(self):
        sm = [r[0] for r in self.rows if r[0] is not None]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "error": str(self.rows[-1][1]) if self.rows else "no samples"}
        bits = 0
        for r in self.rows:
            if r[0] is not None:
                bits |= r[1]
        # NVML clocks-event-reason bits (nvml.h): 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal, 0x40 hw_thermal
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake"}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": self.max_mhz, "reasons": sorted(v for k, v in names.items() if bits & k),
                "samples": len(sm), "how": "NVML, 100 ms period, during the timed region"}


def bind_to_gpu_numa_node(gpu_index):
    """Pins this process to the CPUs NVML reports as local to its GPU, before any pinned host buffer is allocated, so
    that the staging memory of the end-to-end path sits on the GPU's own NUMA node (eight ranks otherwise share
    whatever node the kernel picked and the copies cross the socket link).  Plumbing only; silently skipped if NVML
    or the affinity call is unavailable."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and 64 * w + b < n_cpu]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own code (oracle/_ref, unmodified sources) on all host cores, one worker process
# per core because the reference codecs are not re-entrant; falls back to the oracle port if _ref did not travel.
def _cpu_worker(args):
    kind, blocks, vs, nf, adaptive = args
    import numpy as np
    from oracle import binding
    impl = binding.Ref() if kind == "reference" else binding.Oracle()
    t0 = time.perf_counter()
    out_bytes = 0
    for xs in blocks:
        rc, pay, nb = impl.dega_encode(xs, vs, nf, adaptive)
        assert rc == 0
        rc, y = impl.dega_decode(pay, nb, vs, nf, adaptive, out_cap=xs.size * 4 + 64)
        assert rc == 0 and np.array_equal(y.view(np.uint32), xs.view(np.uint32))
        out_bytes += (nb + 7) // 8
    return time.perf_counter() - t0, out_bytes


def cpu_baseline_run(sample_blocks, cores, kind):
    """sample_blocks: list of float32 arrays (streams).  Returns (GB/s of raw samples for encode+decode, seconds)."""
    import multiprocessing as mp
    shards = [sample_blocks[i::cores] for i in range(cores)]
    shards = [s for s in shards if s]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(len(shards)) as pool:
        res = pool.map(_cpu_worker, [(kind, s, VS, NF, int(ADAPTIVE)) for s in shards])
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)          # excludes interpreter start-up of the workers
    raw = sum(b.size for b in sample_blocks) * 4
    return raw / busy / 1e9, busy, wall


def _cpu_worker_stage(args):
    """Secondary paths: runs one reference chain (list of oracle stage ids) over a list of byte blocks."""
    kind, chain, blocks, vs, nf = args
    from oracle import binding
    impl = binding.Ref() if kind == "reference" else binding.Oracle()
    t0 = time.perf_counter()
    for b in blocks:
        rc, out, nb = impl.run_chain(chain, b, b.size * 8, vs, nf, 0)
        assert rc == 0
    return time.perf_counter() - t0


def cpu_baseline_stage(chain, blocks, cores, kind, vs=32, nf=100.0):
    """GB/s of input bytes for `chain` over `blocks` (uint8 arrays) on `cores` worker processes."""
    import multiprocessing as mp
    shards = [blocks[i::cores] for i in range(cores)]
    shards = [s for s in shards if s]
    with mp.get_context("spawn").Pool(len(shards)) as pool:
        res = pool.map(_cpu_worker_stage, [(kind, chain, s, vs, nf) for s in shards])
    busy = max(res)
    return sum(b.size for b in blocks) / busy / 1e9, busy


def cpu_kind():
    from oracle import binding
    return "reference" if binding.Ref.available() else "port"


def cpu_sample(n_blocks, block_len, seed=0xDC0000):
    """Host copy of the first n_blocks streams of the bench workload (same generator, same seed)."""
    from data_compressor_b200 import synth
    per_channel = (SAMPLES + block_len - 1) // block_len
    n_ch = (n_blocks + per_channel - 1) // per_channel
    x = synth.redd_like(n_ch, SAMPLES, device="cpu", seed=seed).numpy()
    blocks = []
    for c in range(n_ch):
        for b in range(per_channel):
            if len(blocks) < n_blocks:
                blocks.append(x[c, b * block_len:(b + 1) * block_len].copy())
    return blocks


# --------------------------------------------------------------------------------------------------------------
def resolve_block_len(args, ctx=None):
    """--block-len -1 (default) = the two-wave block length of the device in use."""
    if args.block_len >= 0:
        return args.block_len
    try:
        if ctx is None:
            import torch
            if torch.cuda.is_available():
                from data_compressor_b200 import Context
                ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
        if ctx is not None:
            return ctx.two_wave_block_len(args.channels, args.samples)
    except Exception:
        pass
    return BLOCK_LEN


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    args.block_len = resolve_block_len(args)   # same streams as our arm cuts on this box
    from oracle import binding
    binding.build(ref=os.path.isdir("/root/reference/DataCompressor"))
    kind = cpu_kind()
    cores = os.cpu_count() or 1
    # ~2 MB/s per core for encode, ~1.7 MB/s for decode: size one step to a few seconds per core
    blocks_per_core = max(1, int(args.cpu_seconds * 0.9e6 / (args.block_len * 4)))
    blocks = cpu_sample(blocks_per_core * cores, args.block_len)
    for _ in range(args.warmup):
        cpu_baseline_run(blocks[: cores * 4], cores, kind)
    gbs, secs = [], []
    for _ in range(args.steps):
        g, busy, wall = cpu_baseline_run(blocks, cores, kind)
        gbs.append(g)
        secs.append(busy)
    v = sum(b.size for b in blocks) * 4 * len(gbs) / sum(secs) / 1e9
    sample = f"{len(blocks)} streams of {args.block_len} samples ({sum(b.size for b in blocks) * 4 / 1e6:.1f} MB raw) per step, same generator"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32/i64", "data": "synthetic",
        "config": {"workload": workload_name(args.channels, args.samples, args.block_len), "sample": sample},
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from data_compressor_b200 import Context, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: data_compressor_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    ctx = Context(local)
    args.block_len = resolve_block_len(args, ctx)
    C_, S_, L_ = args.channels, args.samples, args.block_len
    # shard = this rank's channels (weak scaling: every GPU gets its own 4096 channels, disjoint seeds)
    x = synth.redd_like(C_, S_, device=dev, seed=0xDC0000 + rank * C_)
    raw_bytes = x.numel() * 4
    n_streams = ctx.n_streams(C_, S_, L_)

    enc = ctx.dega_encode(x, block_len=L_, vs=VS, nf=NF, adaptive=ADAPTIVE)       # allocates outputs + arena once
    y = torch.empty_like(x)
    from data_compressor_b200 import shard

    def step():
        e = ctx.dega_encode(x, block_len=L_, vs=VS, nf=NF, adaptive=ADAPTIVE, out=enc)
        # the only collective on the path: per-shard payload byte counts -> container offsets (16 B per rank).  It is
        # started here, while the GPU is between the two passes, and resolved after the decode: nothing in the decode needs
        # the other ranks' sizes.  Measured side by side on two GPUs (profiles/ubench/xchg_variants.sh): this order 12.74 ms a
        # step, the decode queued first and the collective behind it 12.98 (its kernel then waits for room behind the resident
        # decoder grid), the collective on a stream of its own from preallocated pinned slots 13.1-13.7.
        pending = shard.exchange_sizes_async(e.total_bytes, e.n_streams, dev) if world > 1 else None
        ctx.dega_decode(e, vs=VS, nf=NF, adaptive=ADAPTIVE, out=y)
        if pending is not None:
            pending.result()
        return e

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    ctx.set_profiling(True)
    sampler = ClockSampler(local)
    if not os.environ.get("DC_BENCH_NO_CLOCKS"):
        sampler.start()
    t_wait = time.time()
    while sampler.is_alive() and not sampler.rows and time.time() - t_wait < 5.0:   # NVML initialisation stays out of the timed region
        time.sleep(0.05)
    gc.collect()
    gc.disable()   # a generation-2 collection with torch loaded takes ~50 ms: keep it out of a 14 ms step (as timeit does)
    with torch.cuda.stream(ctx.stream):   # torch's caching allocator keeps one pool per stream: the wrapper's small result tensors
        for _ in range(2):                # must come out of the pool of the stream the timed region runs under, not out of cudaMalloc
            step()                        # (seen as a 10-150 ms first decode call); otherwise exactly what a timed step does
            ctx.kernel_ms()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    for ev in [ev0, ev1] + step_ev:   # torch creates the CUDA event at its first record: do that outside the timed region
        ev.record(ctx.stream)
    barrier()
    launches0 = ctx.launch_count()
    kernel_ms = {}
    pendings = []
    t_enc = t_dec = 0.0
    with torch.cuda.stream(ctx.stream):
        ev0.record(ctx.stream)
        step_ev[0].record(ctx.stream)
        for it in range(args.steps):
            e = ctx.dega_encode(x, block_len=L_, vs=VS, nf=NF, adaptive=ADAPTIVE, out=enc)
            km = ctx.kernel_ms()
            if world > 1:   # started now, resolved when the offsets are needed (here: before the timed region ends)
                pendings.append(shard.exchange_sizes_async(e.total_bytes, e.n_streams, dev))
            ctx.dega_decode(e, vs=VS, nf=NF, adaptive=ADAPTIVE, out=y)
            km2 = ctx.kernel_ms()
            for k, v in list(km.items()) + list(km2.items()):
                kernel_ms[k] = kernel_ms.get(k, 0.0) + v
            step_ev[it + 1].record(ctx.stream)
        sizes = shard.resolve_all(pendings)   # every step's container offsets: inside the timed region, off the steps' critical path
        ev1.record(ctx.stream)
    barrier()
    launches = ctx.launch_count() - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    step_ms = [step_ev[i].elapsed_time(step_ev[i + 1]) for i in range(args.steps)]
    gc.enable()
    sampler.stop()
    ctx.set_profiling(False)
    ok = bool(torch.equal(y.view(torch.int32), x.view(torch.int32))) and not bool(enc.status.any().item())

    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    tot_bytes = torch.tensor([float(enc.total_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_bytes, op=dist.ReduceOp.SUM)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * raw_bytes / (ms_per_step * 1e-3) / 1e9
    kernel_ms = {k: v / args.steps for k, v in kernel_ms.items()}

    # ---- the same step at two other block lengths (not the contract value; kept beside it so that the fitted block
    # length is not the only figure): 1024-sample blocks, and whole channels = the reference's own unit of work
    other = {}
    if not args.no_extra:
        ctx.set_profiling(False)
        for name, bl in (("block_len_1024", 1024), ("whole_channel", 0)):
            if bl == L_:
                continue
            e2 = ctx.dega_encode(x, block_len=bl, vs=VS, nf=NF, adaptive=ADAPTIVE)
            n_rep = 2 if bl == 0 else 5
            ta, tb, tc = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            with torch.cuda.stream(ctx.stream):   # warm up under the stream the timed calls run under (torch's allocator pools are per stream)
                for _ in range(2):
                    ctx.dega_decode(ctx.dega_encode(x, block_len=bl, vs=VS, nf=NF, adaptive=ADAPTIVE, out=e2), vs=VS, nf=NF, adaptive=ADAPTIVE, out=y)
                for ev in (ta, tb, tc):
                    ev.record(ctx.stream)
            barrier()
            with torch.cuda.stream(ctx.stream):
                ta.record(ctx.stream)
                for _ in range(n_rep):
                    ctx.dega_encode(x, block_len=bl, vs=VS, nf=NF, adaptive=ADAPTIVE, out=e2)
                tb.record(ctx.stream)
                for _ in range(n_rep):
                    ctx.dega_decode(e2, vs=VS, nf=NF, adaptive=ADAPTIVE, out=y)
                tc.record(ctx.stream)
            barrier()
            enc_ms2, dec_ms2 = ta.elapsed_time(tb) / n_rep, tb.elapsed_time(tc) / n_rep
            tt = torch.tensor([enc_ms2, dec_ms2], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            enc_ms2, dec_ms2 = float(tt[0].item()), float(tt[1].item())
            ok = ok and bool(torch.equal(y.view(torch.int32), x.view(torch.int32)))
            other[name] = {"block_len": bl, "streams_per_gpu": e2.n_streams, "encode_ms": enc_ms2, "decode_ms": dec_ms2,
                           "value": world * raw_bytes / ((enc_ms2 + dec_ms2) * 1e-3) / 1e9, "encode_gbs": world * raw_bytes / (enc_ms2 * 1e-3) / 1e9,
                           "decode_gbs": world * raw_bytes / (dec_ms2 * 1e-3) / 1e9, "compression_ratio": raw_bytes / e2.total_bytes}
            del e2

    # ---- end to end through the host-buffer C ABI (pinned host memory, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        hx = torch.empty((C_, S_), dtype=torch.float32, pin_memory=True)
        hx.copy_(x)
        torch.cuda.synchronize(dev)
        hxn = hx.numpy()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        kw = dict(vs=VS, nf=NF, adaptive=ADAPTIVE)
        ekw = dict(kw, block_len=L_, payload_cap=int(enc.total_bytes * 1.05) + 4096)   # the compressed size is known from the device run
        ctx2 = ctx   # one context: its job engine runs an encode and a decode side by side and issues the copies of both itself
        for _ in range(2):   # warm-up: the first calls page-lock ~3.3 GB of host buffers and size the device staging
            for slot in (0, 1):
                pay, bl, bo, st = ctx.dega_encode_host(hxn, slot=slot, **ekw)
            yy2, n_out, st2 = ctx2.dega_decode_host(pay, bl, bo, C_, S_, L_, **kw)
        # (a) one context, encode then decode: each call keeps one direction of the link busy
        barrier()
        gc.collect()
        gc.disable()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pay, bl, bo, st = ctx.dega_encode_host(hxn, slot=0, **ekw)
            yy, n_out, st2 = ctx.dega_decode_host(pay, bl, bo, C_, S_, L_, **kw)
        barrier()
        dt_serial = (time.perf_counter() - t0) / e2e_steps
        ok = ok and bool((yy.view("uint32") == hxn.view("uint32")).all())
        # (b) streaming: encodes are queued two deep, the decode of step i's payload is queued as soon as that encode has
        # been collected, and jobs are collected in submission order (dc_wait_next).  The context then always has its next
        # job at hand: an encode (host -> device heavy) runs next to a decode (device -> host heavy), and the tail of one job
        # overlaps the head of the next.  Every step copies every input in and every result out; K encodes and K decodes of
        # every stream in all, pipeline fill and drain inside the timed region.
        n_pipe = max(e2e_steps, args.e2e_stream_steps)

        def stream_steps(k_steps):
            """k_steps encodes + k_steps decodes through the queue; returns (seconds, last decode's results, last encode's results, collect times)."""
            t_start = time.perf_counter()
            pes = {i: ctx.dega_encode_host(hxn, slot=i % 4, wait=False, **ekw) for i in range(min(2, k_steps))}
            pds, prev_, stamps = [], None, []
            for i in range(k_steps):
                prev_ = pes.pop(i).wait()
                stamps.append(time.perf_counter() - t_start)
                pds.append(ctx.dega_decode_host(prev_[0], prev_[1], prev_[2], C_, S_, L_, wait=False, **kw))
                if i + 2 < k_steps:
                    pes[i + 2] = ctx.dega_encode_host(hxn, slot=(i + 2) % 4, wait=False, **ekw)
            for pd in pds:
                last = pd.wait()
            barrier()
            return time.perf_counter() - t_start, last, prev_, stamps

        for slot in (2, 3):   # four sets of encode outputs: encode i+2 may start while the payload of step i-1 is still being decoded
            ctx.dega_encode_host(hxn, slot=slot, **ekw)
        stream_steps(4)       # warm-up of the streaming pattern itself: the second staging block of each kind is allocated here
        yy2.fill(0)           # the decoder's output buffer: the check below must see what the timed steps wrote
        barrier()
        dt_all, (yy, n_out, st2), prev, stamps = stream_steps(n_pipe)
        dt = dt_all / n_pipe
        steady = (stamps[-1] - stamps[len(stamps) // 2]) / (len(stamps) - 1 - len(stamps) // 2) if len(stamps) > 3 else None
        gc.enable()
        ok = ok and bool((yy.view("uint32") == hxn.view("uint32")).all()) and not bool(st2.any()) and not bool(prev[3].any())
        pay = prev[0]
        # (c) what the link itself allows for the same byte counts: pinned copies only, both directions at once
        n_in = raw_bytes + pay.size
        dbuf = torch.empty(n_in, dtype=torch.uint8, device=dev)
        dbuf2 = torch.empty(n_in, dtype=torch.uint8, device=dev)
        # the page-locked buffers the calls above used (no new pinned memory): samples + a payload in, samples + a payload out
        hin = (hx.view(torch.uint8).reshape(-1), ctx._pinned_cache["enc_payload0"][: pay.size])
        hout = (ctx2._pinned_cache["dec_samples"][: raw_bytes], ctx._pinned_cache["enc_payload1"][: pay.size])
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def copies(duplex):
            barrier()
            t = time.perf_counter()
            for _ in range(2):
                with torch.cuda.stream(s1):
                    dbuf[: raw_bytes].copy_(hin[0], non_blocking=True)
                    dbuf[raw_bytes:].copy_(hin[1], non_blocking=True)
                if not duplex:
                    s1.synchronize()
                with torch.cuda.stream(s2):
                    hout[0].copy_(dbuf2[: raw_bytes], non_blocking=True)
                    hout[1].copy_(dbuf2[raw_bytes:], non_blocking=True)
                s1.synchronize()
                s2.synchronize()
            barrier()
            return (time.perf_counter() - t) / 2

        copies(True)
        floor_duplex, floor_serial = copies(True), copies(False)
        del dbuf, dbuf2, hin, hout
        tt = torch.tensor([dt, dt_serial, floor_duplex, floor_serial], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_serial, floor_duplex, floor_serial = (float(v) for v in tt.tolist())
        meta = n_streams * (8 + 8 + 4) + 8
        e2e = {"value": world * raw_bytes / dt / 1e9, "unit": "GB/s", "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": int(raw_bytes + pay.size + meta), "d2h_bytes_per_step": int(raw_bytes + pay.size + 2 * meta),
               "api": "dc_dega_encode_host_async + dc_dega_decode_host_async queued on one context, collected in order with dc_wait_next (include/dc_cuda.h); "
                      "host buffers, every copy inside the timed region (pipeline fill and drain too); the encode of step i runs next to the decode of the "
                      "payload step i-1 produced, so both PCIe directions carry data at once; K encodes + K decodes of every stream in K steps",
               "steps": n_pipe, "steady_state_ms_per_step": steady * 1e3 if steady else None,
               "serial_value": world * raw_bytes / dt_serial / 1e9, "serial_ms_per_step": dt_serial * 1e3,
               "serial_api": "dc_dega_encode_host then dc_dega_decode_host on one context (each call uses one direction of the link)",
               "pcie_floor_ms": floor_duplex * 1e3, "pcie_floor_serial_ms": floor_serial * 1e3,
               "pcie_floor_note": f"pinned cudaMemcpyAsync of the same byte counts ({n_in} B each way), no kernels, measured on every rank at once (max over ranks): "
                                  "both directions together / one after the other",
               "frac_of_pcie_floor": floor_duplex / dt, "host_cpus_bound": numa_cpus}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (4*N read + ceil(bits/8) written per stream)
    peak, peak_src = measured_peak()
    alg_bytes = raw_bytes + enc.total_bytes
    dom = max(("bac_encode", "bac_decode"), key=lambda k: kernel_ms.get(k, 0.0))
    dom_ms = kernel_ms.get(dom, 0.0)
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else None
    traffic = None
    try:   # DRAM bytes per launch from the committed ncu capture, only valid for the shape it was taken on
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if tj["workload"] == {"channels": C_, "samples": S_, "block_len": L_}:
            traffic = tj.get(dom + "_kernel")
    except Exception:
        pass
    issue = None
    try:   # what actually bounds the coder: warp instructions per launch (ncu, same shape) against one per cycle per scheduler
        wi = tj["warp_instructions"][dom + "_kernel"] if traffic is not None else None
        mhz = sampler.summary().get("sm_mhz")
        if wi and mhz and dom_ms > 0:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            peak_i = 4 * sms * mhz * 1e6 / 1e9
            issue = {"achieved": wi / (dom_ms * 1e-3) / 1e9, "peak": peak_i, "unit": "G warp instructions/s", "frac": wi / (dom_ms * 1e-3) / 1e9 / peak_i,
                     "warp_instructions_per_launch": wi, "note": "peak = 4 schedulers x SMs x sampled SM clock, one instruction per cycle; IMAD / LOP3 / SHF / SEL occupy "
                     "their pipe for 2 cycles each on this part (IMAD.HI 4) and the kernels load the ALU pipe and the heavy FMA pipe about "
                     "equally (profiles/r2_ubench_pipes.txt, profiles/r2_coder_experiments.txt); plain adds and predicate logic take 1"}
    except Exception:
        pass
    roofline = {"bound": "hbm", "issue_bound": issue, "kernel": dom + "_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": dom_ms,
                "note": "the coder is sequential per stream and bound by integer instruction issue (1 to 2 cycles per integer instruction per scheduler, profiles/r2_ubench_pipes.txt; DRAM < 4 %): "
                        "DESIGN.md 3.1, profiles/r2d_dega_ncu_summary.txt"}

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            kind = cpu_kind()
            cores = os.cpu_count() or 1
            blocks_per_core = max(1, int(args.cpu_seconds * 0.9e6 / (L_ * 4)))
            blocks = cpu_sample(blocks_per_core * cores, L_)
            g, busy, wall = cpu_baseline_run(blocks, cores, kind)
            cpu = {"value": g, "unit": "GB/s", "cores": cores, "kind": kind, "seconds": busy,
                   "sample": f"{len(blocks)} streams of {L_} samples ({sum(b.size for b in blocks) * 4 / 1e6:.1f} MB raw), encode+decode, "
                             f"one worker process per core, in-process calls (no DCCLI spawn / file I/O)"}
        except Exception as ex:   # the baseline is a reported number, never a reason to lose the GPU line
            cpu = {"value": None, "error": repr(ex)}

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/i64",
        "data": "synthetic",
        "config": {"workload": workload_name(C_, S_, L_), "streams_per_gpu": n_streams, "raw_bytes_per_gpu": raw_bytes,
                   "l2": "inputs (1.4 GB per GPU) are larger than the 126 MB L2; no flush needed", "parallelism": f"channels sharded over {world} GPU(s)"},
        "compression_ratio": world * raw_bytes / float(tot_bytes.item()),
        "encode_ms": sum(kernel_ms.get(k, 0.0) for k in ("front", "bac_encode", "scan", "gather")),
        "decode_ms": sum(kernel_ms.get(k, 0.0) for k in ("bac_decode", "back")),
        "kernel_ms": kernel_ms, "step_ms_min_median_max": [min(step_ms), statistics.median(step_ms), max(step_ms)], "slowest_step": step_ms.index(max(step_ms)), "round_trip_ok": ok,
        "other_block_lengths": other,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------
# Secondary workloads (BASELINE.json configs[2], configs[3]); same timing rules, one JSON line each.
def run_secondary(args):
    import torch
    import torch.distributed as dist
    from data_compressor_b200 import Context, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: data_compressor_b200 has no CPU fallback")
    # weak scaling like the headline workload: every rank runs the same-sized shard (own seed) on its own GPU, there is
    # no exchange on these paths; the job's time is the slowest rank's kernel time
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local)
    ctx.set_profiling(True)
    peak, peak_src = measured_peak()
    steps, warm = args.steps, max(args.warmup, 3)

    def job_ms(enc_ms, dec_ms):
        t = torch.tensor([enc_ms, dec_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item())

    if args.workload == "axdr":
        n_meters, n_samples, vs = args.meters, 96, args.valuesize
        x = synth.meters_15min(n_meters, n_samples, device=dev, seed=0xA0D2 + rank * n_meters)
        enc = ctx.axdr_encode(x, vs=vs)
        y = torch.empty_like(x)
        # 384 MB in + 384 MB out per pass is larger than L2 only in sum: flush L2 between iterations
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        t_enc = t_dec = 0.0
        for it in range(warm + steps):
            flush.zero_()
            enc = ctx.axdr_encode(x, vs=vs, out=enc)
            k1 = ctx.kernel_ms()
            flush.zero_()
            ctx.axdr_decode(enc, vs=vs, out=y)
            k2 = ctx.kernel_ms()
            if it >= warm:
                t_enc += k1["front"]
                t_dec += k2["back"]
        ok = bool(torch.equal(y.view(torch.int32), x.view(torch.int32)))
        raw = x.numel() * 4
        alg = raw + enc.total_bytes
        enc_ms, dec_ms = job_ms(t_enc / steps, t_dec / steps)
        cpu = None
        if rank == 0 and not args.no_cpu:
            # the reference's `encode normalize` on a bounded sample of the same meters, all host cores
            from oracle import binding as ob
            ob.build(ref=os.path.isdir("/root/reference/DataCompressor"))
            kind, cores = cpu_kind(), os.cpu_count() or 1
            per_core = max(1, int(args.cpu_seconds * 15e6 / (n_samples * 4)))          # ~18 MB/s per core
            xs = x[: per_core * cores].cpu().numpy()
            gbs, busy = cpu_baseline_stage([ob.NORMALIZE], [row.view("uint8") for row in xs], cores, kind, vs, NF)
            cpu = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind, "seconds": busy,
                   "sample": f"{xs.shape[0]} meters of {n_samples} samples ({xs.nbytes / 1e6:.1f} MB raw), encode normalize, one worker process per core"}
        line = {"metric": "A-XDR (normalize only) encode GB/s of raw samples", "value": world * raw / (enc_ms * 1e-3) / 1e9, "unit": "GB/s",
                "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": enc_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32->i64", "data": "synthetic",
                "config": {"workload": f"configs[2]: A-XDR (normalize only) over {n_meters} synthetic meters x {n_samples} samples (15 min) per GPU, valuesize {vs}",
                           "l2": "256 MiB flush write between iterations"},
                "decode_gbs": world * raw / (dec_ms * 1e-3) / 1e9, "decode_ms": dec_ms, "round_trip_ok": ok,
                "roofline": {"bound": "hbm", "kernel": "axdr_encode_kernel", "achieved": alg / (enc_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": alg / (enc_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": int(alg)},
                "cpu_baseline": cpu, "gpu_launches": ctx.launch_count()}
    else:   # lzmh
        n_blocks, block = args.blocks, 65536
        chans = (n_blocks * block + 2 * 86400 - 1) // (2 * 86400)
        data = synth.lzmh_blocks(n_blocks, block, device=dev, seed=0xDC0000 + rank * chans)
        t_enc = t_dec = 0.0
        for it in range(warm + steps):
            payload, bit_len, byte_off, status, total, in_off = ctx.lzmh_encode(data, block_bytes=block)
            k1 = ctx.kernel_ms()
            out, out_len, st2 = ctx.lzmh_decode(payload, total, bit_len, byte_off, in_off)
            k2 = ctx.kernel_ms()
            if it >= warm:
                t_enc += k1["lzmh_encode"] + k1["scan"] + k1["gather"]
                t_dec += k2["lzmh_decode"]
        raw = data.numel()
        # blocks ending in an un-ranked zero byte lose it in the reference decoder (lzmh.c:571): compare lengths too
        same = int((out_len == (in_off[1:] - in_off[:-1])).sum().item())
        ok = bool(torch.equal(out[: raw], data)) if same == n_blocks else None
        enc_ms, dec_ms = job_ms(t_enc / steps, t_dec / steps)
        alg = raw + total
        cpu = None
        if rank == 0 and not args.no_cpu:
            # the reference's `encode lzmh` on a bounded sample of the same blocks, all host cores
            from oracle import binding as ob
            ob.build(ref=os.path.isdir("/root/reference/DataCompressor"))
            kind, cores = cpu_kind(), os.cpu_count() or 1
            per_core = max(1, int(args.cpu_seconds * 2.5e6 / block))                       # ~2.7 MB/s per core
            hb = data[: per_core * cores * block].cpu().numpy().reshape(-1, block)
            gbs, busy = cpu_baseline_stage([ob.LZMH_ENC], [row for row in hb], cores, kind)
            cpu = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind, "seconds": busy,
                   "sample": f"{hb.shape[0]} blocks of {block} bytes ({hb.nbytes / 1e6:.1f} MB), encode lzmh, one worker process per core"}
        line = {"metric": "LZMH encode GB/s of input bytes", "value": world * raw / (enc_ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world, "steps": steps,
                "warmup": warm, "ms_per_step": enc_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": {"workload": f"configs[3]: LZMH over {n_blocks} blocks of {block} bytes per GPU (synthetic 16-bit integer-watt load profiles)",
                           "l2": "input larger than L2" if raw > (126 << 20) else "input fits L2"},
                "decode_gbs": world * raw / (dec_ms * 1e-3) / 1e9, "decode_ms": dec_ms, "compression_ratio": raw / total,
                "blocks_fully_restored": same, "round_trip_ok": ok,
                "roofline": {"bound": "hbm", "kernel": "lzmh_encode_kernel", "achieved": alg / (enc_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": alg / (enc_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": int(alg), "note": "sequential greedy parse per block: issue/latency bound"},
                "cpu_baseline": cpu, "gpu_launches": ctx.launch_count()}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: granularity sweep (1 s / 1 min / 15 min) with block-size and value-size variations.
# One JSON line per point (kernel-pipeline times, same rules as the headline line); the coarser series are derived
# on the device with dc_aggregate_device (sums of 60 / 900 one-second samples, turned into mean power and re-quantised to
# two decimals so that the reference round trip stays bit-exact: sums themselves leave float32's exact range).
def run_sweep(args):
    import torch
    import torch.distributed as dist
    from data_compressor_b200 import Context, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: data_compressor_b200 has no CPU fallback")
    # weak scaling like the headline workload: every rank sweeps its own 4096 channels (own seed) on its own GPU, a
    # point's time is the slowest rank's; rank 0 prints, and times the reference on the host cores once per granularity
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local)
    ctx.set_profiling(True)
    steps, warm = max(1, min(args.steps, 5)), 2
    x1 = synth.redd_like(args.channels, SAMPLES, device=dev, seed=0xDC0000 + rank * args.channels)
    hundred = torch.tensor(100.0, dtype=torch.float32, device=dev)
    series = {"1s": x1}
    for name, k in (("1min", 60), ("15min", 900)):
        a = ctx.aggregate(x1, k) / float(k)                                   # mean power of the interval
        series[name] = (torch.round(a * hundred) / hundred).contiguous()     # two decimals again: exact reference round trip
    kind, cores = None, os.cpu_count() or 1
    if rank == 0 and not args.no_cpu:
        from oracle import binding as ob
        ob.build(ref=os.path.isdir("/root/reference/DataCompressor"))
        kind = cpu_kind()
    for gran, x in series.items():
        n = x.shape[1]
        cpu = None
        if kind is not None:
            # the reference on whole series of this granularity (one plain DCCLI-style run per channel), a bounded sample
            per_core = max(1, int(args.cpu_seconds * 0.9e6 / (n * 4)))
            rows = x[: min(x.shape[0], per_core * cores)].cpu().numpy()
            g, busy, wall = cpu_baseline_run([r.copy() for r in rows], cores, kind)
            cpu = {"value": g, "unit": "GB/s", "cores": cores, "kind": kind, "seconds": busy,
                   "sample": f"{rows.shape[0]} whole channels of {n} samples ({rows.nbytes / 1e6:.1f} MB raw), valuesize 32, encode+decode, one worker process per core"}
        for vs in (17, 24, 32, 64):
            ok_vs = torch.tensor([float(x.max().item()) * 100.0 + 0.5 <= float(2 ** (vs - 1) - 1)], dtype=torch.int32, device=dev)
            if world > 1:
                dist.all_reduce(ok_vs, op=dist.ReduceOp.MIN)
            if not bool(ok_vs.item()):
                continue      # normalize.c:21-22 would reject the series at this value size (on some rank)
            for bl in (0, 16384, 4096, 1024):
                if bl >= n and bl != 0:
                    continue
                t_enc = t_dec = 0.0
                enc = None
                for it in range(warm + steps):
                    enc = ctx.dega_encode(x, block_len=bl, vs=vs, nf=NF, adaptive=ADAPTIVE, out=enc)
                    k1 = ctx.kernel_ms()
                    y, n_out, status = ctx.dega_decode(enc, vs=vs, nf=NF, adaptive=ADAPTIVE)
                    k2 = ctx.kernel_ms()
                    if it >= warm:
                        t_enc += sum(k1[k] for k in ("front", "bac_encode", "scan", "gather"))
                        t_dec += sum(k2[k] for k in ("bac_decode", "back"))
                ok = bool(torch.equal(y.view(torch.int32), x.view(torch.int32))) and not bool(enc.status.any().item())
                raw = x.numel() * 4
                t = torch.tensor([t_enc / steps, t_dec / steps, -float(ok)], dtype=torch.float64, device=dev)
                tot = torch.tensor([float(enc.total_bytes)], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
                enc_ms, dec_ms, ok = float(t[0].item()), float(t[1].item()), float(t[2].item()) == -1.0
                if rank == 0:
                    line = {
                        "metric": "DEGA encode / decode GB/s of raw samples (configs[4] sweep)", "granularity": gran, "samples_per_channel": n,
                        "channels_per_gpu": args.channels, "valuesize": vs, "block_len": bl, "streams_per_gpu": enc.n_streams, "encode_ms": enc_ms,
                        "decode_ms": dec_ms, "encode_gbs": world * raw / (enc_ms * 1e-3) / 1e9, "decode_gbs": world * raw / (dec_ms * 1e-3) / 1e9,
                        "value": world * raw / ((enc_ms + dec_ms) * 1e-3) / 1e9, "unit": "GB/s",
                        "compression_ratio": world * raw / float(tot.item()), "round_trip_ok": ok, "n_gpus": world, "steps": steps, "scaling": "weak",
                        "data": "synthetic"}
                    if vs == 32 and bl == 0:
                        line["cpu_baseline"] = cpu
                    print(json.dumps(line), flush=True)
                del y
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------
# the csv codec through the host-buffer calls (text in pinned host memory, copies inside the timed region);
# the reference's ReadCSV / WriteCSV (one loop each: one host core) timed on a bounded sample of the same text.
def _csv_ref_worker(arg):
    kind, text = arg
    from oracle import binding as ob
    impl = ob.Ref() if kind == "reference" else ob.Oracle()
    t0 = time.perf_counter()
    v = impl.csv_decode(text)
    t1 = time.perf_counter()
    back = impl.csv_encode(v)
    t2 = time.perf_counter()
    return v.size, t1 - t0, t2 - t1, back == text


def run_csv(args):
    import ctypes as C
    import multiprocessing as mp
    import numpy as np
    import torch
    from data_compressor_b200 import Context, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: data_compressor_b200 has no CPU fallback")
    ctx = Context(0)
    n_ch = max(1, args.csv_lines // SAMPLES)
    x = synth.redd_like(n_ch, SAMPLES, device="cuda").cpu().numpy().reshape(-1)     # two-decimal watt readings, like DCCLI/testdata/input.txt
    text = ctx.csv_encode(x)                                                         # the text itself comes from the library (checked below)
    n_lines = x.size
    pin_t = torch.empty(len(text), dtype=torch.uint8).pin_memory()
    pin_t.numpy()[:] = np.frombuffer(text, dtype=np.uint8)
    pin_v = torch.empty(n_lines, dtype=torch.float32).pin_memory()
    pin_o = torch.empty(len(text) + 64, dtype=torch.uint8).pin_memory()
    n, m = C.c_uint64(0), C.c_uint64(0)
    steps, warm = max(1, min(args.steps, 10)), 3
    t_dec = t_enc = 0.0
    launches0 = ctx.launch_count()
    gc.collect()
    gc.disable()
    for it in range(warm + steps):
        t0 = time.perf_counter()
        rc = ctx.lib.dc_csv_decode_host(ctx.h, pin_t.data_ptr(), len(text), 1, b",", pin_v.data_ptr(), n_lines, C.byref(n))
        t1 = time.perf_counter()
        rc2 = ctx.lib.dc_csv_encode_host(ctx.h, pin_v.data_ptr(), n_lines, 1, b",", 2, pin_o.data_ptr(), pin_o.numel(), C.byref(m))
        t2 = time.perf_counter()
        assert rc == 0 and rc2 == 0 and n.value == n_lines and m.value == len(text)
        if it >= warm:
            t_dec += t1 - t0
            t_enc += t2 - t1
    gc.enable()
    launches = (ctx.launch_count() - launches0) // (warm + steps) * steps
    ok = bool(np.array_equal(pin_v.numpy().view(np.uint32), x.view(np.uint32))) and pin_o.numpy()[: len(text)].tobytes() == text
    dec_s, enc_s = t_dec / steps, t_enc / steps
    cpu = None
    if not args.no_cpu:
        from oracle import binding as ob
        ob.build(ref=os.path.isdir("/root/reference/DataCompressor"))
        kind = cpu_kind()
        sample_lines = min(n_lines, int(args.cpu_seconds * 1.0e6))                   # ~2 M lines/s each way on one core
        cut = text.rfind(b"\n", 0, min(len(text), sample_lines * 6)) + 1            # ~6 bytes per line
        with mp.get_context("spawn").Pool(1) as pool:
            lines, ds, es, same = pool.map(_csv_ref_worker, [(kind, text[:cut])])[0]
        cpu = {"value": lines / ds / 1e6, "unit": "M lines/s", "cores": 1, "kind": kind, "encode_value": lines / es / 1e6, "round_trip_ok": same,
               "sample": f"{lines} lines ({cut / 1e6:.1f} MB of text), decode csv then encode csv, one core (ReadCSV and WriteCSV are single loops)"}
    peak, peak_src = measured_peak()
    alg = len(text) + 4 * n_lines
    print(json.dumps({
        "metric": "csv decode M lines/s through the host-buffer call", "value": n_lines / dec_s / 1e6, "unit": "M lines/s", "n_gpus": 1, "steps": steps,
        "warmup": warm, "ms_per_step": dec_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8->f32", "data": "synthetic",
        "config": {"workload": f"SURVEY 8f-3: {n_lines} lines of two-decimal watt readings ({len(text) / 1e6:.1f} MB of text), column 1, pinned host buffers"},
        "encode_value": n_lines / enc_s / 1e6, "encode_ms": enc_s * 1e3, "text_gbs": len(text) / dec_s / 1e9, "round_trip_ok": ok,
        "e2e": {"value": n_lines / dec_s / 1e6, "unit": "M lines/s", "h2d_bytes_per_step": len(text), "d2h_bytes_per_step": 4 * n_lines},
        "roofline": {"bound": "pcie", "achieved": alg / dec_s / 1e9, "unit": "GB/s", "peak": None, "frac": None, "traffic": None,
                     "note": "the call is two PCIe copies (text in, floats out) around ~1.5 ms of kernels: profiles/r1_csv_codec.txt"},
        "cpu_baseline": cpu, "gpu_launches": launches}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=CHANNELS)
    ap.add_argument("--samples", type=int, default=SAMPLES)
    ap.add_argument("--block-len", type=int, default=-1, help="samples per resumable block; 0 = whole channels; -1 = the device's two-wave length")
    ap.add_argument("--no-extra", action="store_true", help="skip the 1024-sample-block and whole-channel figures")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="target CPU-baseline work per core")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-stream-steps", type=int, default=20, help="steps of the streaming end-to-end measurement (fill and drain are inside the timed region)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="dega", choices=["dega", "axdr", "lzmh", "sweep", "csv"],
                    help="dega = the contract line; axdr/lzmh = secondary paths; sweep = configs[4]; csv = the text codec (SURVEY 8f-3)")
    ap.add_argument("--csv-lines", type=int, default=50_000_000)
    ap.add_argument("--meters", type=int, default=1_000_000)
    ap.add_argument("--valuesize", type=int, default=32)
    ap.add_argument("--blocks", type=int, default=16384)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "sweep":
        run_sweep(args)
    elif args.workload == "csv":
        run_csv(args)
    elif args.workload != "dega":
        run_secondary(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
