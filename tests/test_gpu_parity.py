"""Parity tests proper: the sm_100a kernels, called through the C ABI (include/dc_cuda.h), against the oracle and
the committed golden vectors.  Bit-exact everywhere (integer/byte work; the two float32 steps use IEEE
round-to-nearest intrinsics and are compared bit for bit as well)."""
import numpy as np
import pytest
import torch

from conftest import split_streams
from oracle import binding as ob

pytestmark = pytest.mark.gpu

DEGA_ENC, DEGA_DEC = 10, 11


def profile(rng, shape, step=200, base=30000):
    k = np.abs(np.cumsum(rng.integers(-step, step + 1, shape), axis=-1) + base)
    return (k.astype(np.float32) / np.float32(100.0)).astype(np.float32)


# ---- plugin face: one stage, one stream ----------------------------------------------------------------------
def test_golden_vectors_stage_by_stage(ctx, golden):
    """Every golden chain, one GPU stage at a time (what DCCLI's stage loop would do with GPU plugins)."""
    for name in golden.names:
        g = golden.get(name)
        if any(s in (ob.LZMH_ENC, ob.LZMH_DEC) for s in g["chain"]):
            continue
        data, nbits, rc = g["input"].tobytes(), g["in_bits"], 0
        for st in g["chain"]:
            rc, data, nbits = ctx.stage_run(st, data, nbits, g["vs"], g["nf"], g["adaptive"])
            if rc != 0:
                break
        assert rc == g["rc"], name
        if rc == 0:
            assert nbits == g["out_bits"], name
            if g["out"].size:
                assert data == g["out"].tobytes(), name


def test_make_test_on_gpu(ctx, golden):
    """The reference's `make test`: DEGA round trip of testdata/input.txt, plus the exact compressed bytes."""
    g = golden.get("testfile_stage4")
    rc, out, nb = ctx.stage_run(DEGA_ENC, golden.testfile, None, 32, 100.0, True)
    assert rc == 0 and nb == 1354613 and out == g["out"].tobytes()
    rc, dec, nb2 = ctx.stage_run(DEGA_DEC, out, nb, 32, 100.0, True)
    assert rc == 0 and nb2 == 3200000 and dec == golden.testfile.tobytes()


@pytest.mark.parametrize("seed", range(6))
def test_random_stage_chains_vs_oracle(ctx, oracle, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(0, 4000))
    x = profile(rng, n, step=int(rng.choice([3, 300, 30000])))
    vs = int(rng.choice([22, 24, 32, 33, 48, 64]))
    ad = int(rng.integers(0, 2))
    data, nbits = x.tobytes(), n * 32
    for st in ob.DEGA_ENC + ob.DEGA_DEC:
        want = oracle.run_chain([st], data, nbits, vs, 100.0, ad)
        got = ctx.stage_run(st, data, nbits, vs, 100.0, ad)
        assert got == want, (st, vs, ad, n)
        data, nbits = got[1], got[2]


@pytest.mark.parametrize("seed", range(4))
def test_bac_random_bits(ctx, oracle, seed):
    rng = np.random.default_rng(50 + seed)
    n = int(rng.choice([0, 1, 33, 1000, 20000, 70000]))
    p1 = float(rng.choice([0.001, 0.1, 0.5, 0.97]))
    bits = np.packbits((rng.random(n) < p1).astype(np.uint8))
    for ad in (0, 1):
        want = oracle.run_chain([ob.BAC_ENC], bits, n, adaptive=ad)
        got = ctx.stage_run(ob.BAC_ENC, bits, n, adaptive=ad)
        assert got == want
        assert ctx.stage_run(ob.BAC_DEC, got[1], got[2], adaptive=ad) == oracle.run_chain([ob.BAC_DEC], got[1], got[2], adaptive=ad)


def test_bac_long_pending_runs(ctx, oracle):
    """Runs of middle-third symbols pile up E3 steps (pending > 16 leaves the branch-free emitter)."""
    for n_ones in (5, 12, 20, 40, 1000):
        for tail in ([], [0], [0, 1, 1, 0]):
            bits = np.array([1] * n_ones + tail, dtype=np.uint8)
            packed = np.packbits(bits)
            for ad in (0, 1):
                want = oracle.run_chain([ob.BAC_ENC], packed, bits.size, adaptive=ad)
                got = ctx.stage_run(ob.BAC_ENC, packed, bits.size, adaptive=ad)
                assert got == want, (n_ones, tail, ad)
                back = ctx.stage_run(ob.BAC_DEC, got[1], got[2], adaptive=ad)
                assert back[0] == 0 and back[2] == bits.size and back[1] == packed.tobytes()


@pytest.mark.parametrize("adaptive", [0, 1])
def test_bac_carry_through_stored_words(ctx, oracle, adaptive):
    """Pending runs of up to 300 bits resolved with and without a carry: the carry has to ripple through the encoder's
    accumulators and through words it has already stored (CarryWriter / bac_ripple in bac.cuh)."""
    from conftest import straddling_bits
    rng = np.random.default_rng(3 + adaptive)
    for n in (1, 31, 32, 33, 63, 64, 65, 129, 300):
        for tail in ([], [0], [1, 1, 1], list(rng.integers(0, 2, 40))):
            for prefix, resolve in (((), "up"), ((), "down"), (tuple(rng.integers(0, 2, int(rng.integers(1, 70)))), None)):
                bits = straddling_bits(n, tail, adaptive, prefix, resolve)
                packed = np.packbits(bits)
                want = oracle.run_chain([ob.BAC_ENC], packed, bits.size, adaptive=adaptive)
                got = ctx.stage_run(ob.BAC_ENC, packed, bits.size, adaptive=adaptive)
                assert got == want, (n, tail, prefix, resolve)
                back = ctx.stage_run(ob.BAC_DEC, got[1], got[2], adaptive=adaptive)
                assert back[0] == 0 and back[2] == bits.size and back[1] == packed.tobytes()


def test_bac_groups_that_run_out_of_room_are_redone(ctx, oracle):
    """After 60000 zeros a one costs ~13 bits: the unrolled 16-symbol groups overflow their 31 / 32 bits of room and are
    redone symbol by symbol (encoder and decoder); > 20 halvings on the way; balanced stretches keep the symbols
    trading places (the general, swap-aware group form)."""
    rng = np.random.default_rng(7)
    bits = np.concatenate([np.zeros(60000, np.uint8), (rng.random(80000) < 0.5).astype(np.uint8), np.ones(60000, np.uint8),
                           np.tile(np.array([0, 1], np.uint8), 5000), (rng.random(30000) < 0.02).astype(np.uint8)])
    packed = np.packbits(bits)
    for ad in (1, 0):
        want = oracle.run_chain([ob.BAC_ENC], packed, bits.size, adaptive=ad)
        got = ctx.stage_run(ob.BAC_ENC, packed, bits.size, adaptive=ad)
        assert got == want
        back = ctx.stage_run(ob.BAC_DEC, got[1], got[2], adaptive=ad)
        assert back[0] == 0 and back[2] == bits.size and back[1] == packed.tobytes()


def test_warp_interleaved_scratch_with_ragged_stream_counts(ctx, oracle):
    """Stream counts that are not multiples of 32 (partial warps in the interleaved SEG scratch) and streams of very
    different lengths inside one warp."""
    import torch
    rng = np.random.default_rng(21)
    for nc, cl, bl in ((1, 37, 0), (3, 100, 7), (33, 65, 0), (5, 1000, 33)):
        k = np.abs(np.cumsum(rng.integers(-3000, 3001, (nc, cl)), axis=1) + 50000)
        x = (k.astype(np.float32) / np.float32(100.0)).astype(np.float32)
        enc = ctx.dega_encode(torch.from_numpy(x).cuda(), block_len=bl)
        assert not enc.status.any().item()
        for s_, xs in enumerate(split_streams(x, bl)):
            rc, want, nb = oracle.dega_encode(xs)
            assert rc == 0 and nb == int(enc.bit_len[s_]) and enc.stream_bytes(s_) == want, (nc, cl, bl, s_)
        y, n_out, status = ctx.dega_decode(enc)
        assert not status.any().item() and np.array_equal(y.cpu().numpy().view(np.uint32), x.view(np.uint32))


def test_garbage_and_truncated_inputs(ctx, oracle):
    rng = np.random.default_rng(9)
    junk = rng.integers(0, 256, 300, dtype=np.uint8)
    for nbits in (1, 15, 16, 17, 333, 2400):
        for st, kw in ((ob.SEG_DEC, dict(vs=32)), (ob.SEG_DEC, dict(vs=8)), (ob.BAC_DEC, dict(adaptive=1)), (ob.BAC_DEC, dict(adaptive=0))):
            a = ctx.stage_run(st, junk, nbits, **kw)
            b = oracle.run_chain([st], junk, nbits, **kw)
            assert a[0] == b[0] and (a[0] != 0 or a == b), (st, nbits, kw)


# ---- batch face --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("block_len", [0, 1, 7, 64, 1000])
@pytest.mark.parametrize("adaptive", [True, False])
def test_batch_encode_matches_oracle_per_stream(ctx, oracle, block_len, adaptive):
    rng = np.random.default_rng(block_len + 17)
    nc, cl = 37, 333
    x = profile(rng, (nc, cl))
    enc = ctx.dega_encode(torch.from_numpy(x).cuda(), block_len=block_len, adaptive=adaptive)
    bit_len = enc.bit_len.cpu().numpy()
    byte_off = enc.byte_off.cpu().numpy()
    payload = enc.payload[: enc.total_bytes].cpu().numpy()
    assert not enc.status.any().item()
    streams = split_streams(x, block_len)
    assert len(streams) == bit_len.size
    for s, xs in enumerate(streams):
        rc, ob_, nb = oracle.dega_encode(xs, adaptive=int(adaptive))
        assert rc == 0 and nb == bit_len[s], s
        o = int(byte_off[s])
        assert payload[o:o + len(ob_)].tobytes() == ob_, s
        assert int(byte_off[s + 1]) - o == len(ob_)
    assert int(byte_off[-1]) == enc.total_bytes
    y, n_out, status = ctx.dega_decode(enc, adaptive=adaptive)
    assert not status.any().item()
    assert torch.equal(y.view(torch.int32).cpu(), torch.from_numpy(x).view(torch.int32))
    assert n_out.cpu().tolist() == [len(s) for s in streams]


@pytest.mark.parametrize("vs", [17, 24, 40, 64])
def test_batch_value_sizes(ctx, oracle, vs):
    rng = np.random.default_rng(vs)
    x = profile(rng, (9, 500), step=50, base=2000)
    enc = ctx.dega_encode(torch.from_numpy(x).cuda(), block_len=128, vs=vs)
    streams = split_streams(x, 128)
    for s, xs in enumerate(streams):
        assert enc.stream_bytes(s) == oracle.dega_encode(xs, vs=vs)[1]
    y, n_out, status = ctx.dega_decode(enc, vs=vs)
    assert torch.equal(y.view(torch.int32).cpu(), torch.from_numpy(x).view(torch.int32))


def test_batch_streams_with_long_code_words(ctx, oracle):
    """Deltas of 2^15 and more (SEG code words of 33 .. 63 bits: appliances switching) between short ones, at every
    value size the specialised back end takes; the decoder's medium path and the general parser must agree with the oracle."""
    rng = np.random.default_rng(71)
    nc, cl = 40, 700
    k = np.abs(np.cumsum(rng.integers(-200, 201, (nc, cl)), axis=1) + 30000)
    jumps = rng.random((nc, cl)) < 0.03
    k = np.abs(k + np.cumsum(np.where(jumps, rng.integers(-1_000_000, 1_000_001, (nc, cl)), 0), axis=1)) % (1 << 21)   # float32 round trip stays exact (SURVEY a-2)
    x = (k.astype(np.float32) / np.float32(100.0)).astype(np.float32)
    for vs in (25, 32, 40):
        enc = ctx.dega_encode(torch.from_numpy(x).cuda(), block_len=175, vs=vs)
        assert not enc.status.any().item()
        for s, xs in enumerate(split_streams(x, 175)):
            if s % 7 == 0:
                assert enc.stream_bytes(s) == oracle.dega_encode(xs, vs=vs)[1], (vs, s)
        y, n_out, status = ctx.dega_decode(enc, vs=vs)
        assert not status.any().item() and torch.equal(y.view(torch.int32).cpu(), torch.from_numpy(x).view(torch.int32)), vs
        # the decoder against the oracle's decoder on the same payload (not only a round trip)
        for s in (0, 33, 159):
            pay = enc.stream_bytes(s)
            rc, want = oracle.dega_decode(pay, int(enc.bit_len[s]), vs=vs)
            assert rc == 0 and np.array_equal(want.view(np.uint32), y.cpu().numpy().reshape(-1, 175)[s].view(np.uint32))


def test_batch_failed_streams_are_isolated(ctx, oracle):
    x = np.full((4, 50), 12.5, dtype=np.float32)
    x[1, 20] = -3.0            # diff rejects
    x[3, 0] = 3.0e9            # normalize rejects (a-1)
    enc = ctx.dega_encode(torch.from_numpy(x).cuda())
    assert enc.status.cpu().tolist() == [0, -1, 0, -1]
    assert enc.bit_len.cpu().tolist()[1] == 0 and enc.bit_len.cpu().tolist()[3] == 0
    assert enc.stream_bytes(0) == oracle.dega_encode(x[0])[1]
    assert enc.stream_bytes(2) == oracle.dega_encode(x[2])[1]


def test_empty_and_tiny_batches(ctx, oracle):
    enc = ctx.dega_encode(torch.zeros((3, 0), dtype=torch.float32, device="cuda"))
    assert enc.bit_len.cpu().tolist() == [3, 3, 3] and enc.total_bytes == 3          # empty stream -> 3 bits `001`
    assert enc.stream_bytes(1) == bytes([0x20])
    enc = ctx.dega_encode(torch.tensor([[1.0, 2.0, 1.5]], device="cuda"))
    assert enc.stream_bytes(0) == bytes.fromhex("f9c008475e84") and int(enc.bit_len[0]) == 47


def test_host_buffer_face(ctx, oracle):
    rng = np.random.default_rng(3)
    x = profile(rng, (11, 700))
    payload, bit_len, byte_off, status = ctx.dega_encode_host(x, block_len=256)
    streams = split_streams(x, 256)
    for s, xs in enumerate(streams):
        o = int(byte_off[s])
        want = oracle.dega_encode(xs)[1]
        assert payload[o:o + len(want)].tobytes() == want
    y, n_out, status = ctx.dega_decode_host(payload, bit_len, byte_off, 11, 700, 256)
    assert np.array_equal(y.view(np.uint32), x.view(np.uint32)) and not status.any()


def test_host_buffer_face_chunked_overlap(ctx, oracle):
    """Enough streams for the host-buffer calls to split into overlapped channel chunks (>= 2 waves of coder CTAs)."""
    from data_compressor_b200 import synth
    nc, cl, bl = 1301, 4000, 16            # 325,250 streams -> 2 chunks of 651/650 channels
    x = synth.redd_like(nc, cl, device="cpu", seed=77).numpy()
    payload, bit_len, byte_off, status = ctx.dega_encode_host(x, block_len=bl)
    assert not status.any() and int(byte_off[-1]) == payload.size
    assert np.array_equal(byte_off[1:] - byte_off[:-1], (bit_len + 7) // 8)
    rng = np.random.default_rng(1)
    for s in list(rng.integers(0, bit_len.size, 12)) + [0, 651 * 250 - 1, 651 * 250, bit_len.size - 1]:
        c, b = divmod(int(s), 250)
        want = oracle.dega_encode(x[c, b * bl:(b + 1) * bl])[1]
        o = int(byte_off[s])
        assert payload[o:o + len(want)].tobytes() == want, s
    y, n_out, st2 = ctx.dega_decode_host(payload, bit_len, byte_off, nc, cl, bl)
    assert np.array_equal(y.view(np.uint32), x.view(np.uint32)) and not st2.any() and (n_out == bl).all()


def test_axdr_normalize_only(ctx, oracle):
    rng = np.random.default_rng(5)
    x = profile(rng, (50, 96), step=40, base=500)
    for vs in (17, 24, 32):
        enc = ctx.axdr_encode(torch.from_numpy(x).cuda(), vs=vs)
        for s in (0, 7, 49):
            want = oracle.run_chain([ob.NORMALIZE], x[s], 96 * 32, vs)
            assert int(enc.bit_len[s]) == want[2] and enc.stream_bytes(s) == want[1]
        y, n_out, status = ctx.axdr_decode(enc, vs=vs)
        assert torch.equal(y.view(torch.int32).cpu(), torch.from_numpy(x).view(torch.int32))


def test_axdr_failed_meter_falls_back_to_dense_layout(ctx, oracle):
    rng = np.random.default_rng(6)
    x = profile(rng, (40, 96), step=40, base=500)
    x[7, 13] = 5.0e9                       # normalize rejects this meter (normalize.c:21-22)
    enc = ctx.axdr_encode(torch.from_numpy(x).cuda(), vs=32)
    st = enc.status.cpu().tolist()
    assert st[7] == -1 and sum(1 for v in st if v) == 1 and int(enc.bit_len[7]) == 0
    assert enc.total_bytes == 39 * 96 * 4
    for s in (0, 6, 8, 39):
        assert enc.stream_bytes(s) == oracle.run_chain([ob.NORMALIZE], x[s], 96 * 32, 32)[1]


def test_axdr_million_meters_round_trip(ctx, oracle):
    """BASELINE.json config 3 at full size: 1,000,000 meters x 96 samples, normalize only."""
    from data_compressor_b200 import synth
    x = synth.meters_15min(1_000_000, 96, device="cuda")
    for vs in (32, 24):
        enc = ctx.axdr_encode(x, vs=vs)
        assert enc.total_bytes == 1_000_000 * 96 * vs // 8 and not enc.status.any().item()
        y, n_out, status = ctx.axdr_decode(enc, vs=vs)
        assert torch.equal(y.view(torch.int32), x.view(torch.int32))
        for s in (0, 123_456, 999_999):
            assert enc.stream_bytes(s) == oracle.run_chain([ob.NORMALIZE], x[s].cpu().numpy(), 96 * 32, vs)[1]


# ---- LZMH ------------------------------------------------------------------------------------------------------
def test_lzmh_golden_vectors(ctx, golden):
    for name in golden.names:
        g = golden.get(name)
        if g["chain"] not in ([ob.LZMH_ENC], [ob.LZMH_DEC]):
            continue
        rc, out, nb = ctx.stage_run(g["chain"][0], g["input"].tobytes(), g["in_bits"])
        assert rc == g["rc"], name
        assert nb == g["out_bits"], name
        import hashlib
        assert hashlib.sha256(out).hexdigest() == g["sha256"], name


@pytest.mark.parametrize("seed", range(6))
def test_lzmh_random_blocks_vs_oracle(ctx, oracle, seed):
    rng = np.random.default_rng(400 + seed)
    for n in (0, 1, 2, 3, 127, 128, 129, 274, 275, 402, 403, 404, 1000, 1023, 1024, 1025, 5000, 70000):
        kind = (seed + n) % 3
        if kind == 0:
            data = rng.integers(0, 4, n, dtype=np.uint8)                 # long matches
        elif kind == 1:
            data = np.frombuffer(("".join(f"{v:.2f}\n" for v in rng.random(n // 5 + 1) * 300))[:n].encode(), dtype=np.uint8)
        else:
            data = rng.integers(0, 256, n, dtype=np.uint8)               # incompressible: > 48 distinct literals
        want = oracle.run_chain([ob.LZMH_ENC], data, n * 8)
        got = ctx.stage_run(ob.LZMH_ENC, data, n * 8)
        assert got == want, (n, kind)
        assert ctx.stage_run(ob.LZMH_DEC, got[1], got[2]) == oracle.run_chain([ob.LZMH_DEC], got[1], got[2]), (n, kind)


def test_lzmh_decoder_on_arbitrary_bits(ctx, oracle):
    """Streams no encoder wrote: random bits (every token kind, zero offsets out of the empty recent list, copies from
    history that was never written, prefix codes that match nothing, tokens cut off by the end) and overlapping copies
    of every period 1..40."""
    rng = np.random.default_rng(77)
    for n_bytes in (1, 2, 3, 8, 64, 300, 2000):
        for _ in range(12):
            junk = rng.integers(0, 256, n_bytes, dtype=np.uint8)
            for nbits in (n_bytes * 8, n_bytes * 8 - 3):
                assert ctx.stage_run(ob.LZMH_DEC, junk, nbits) == oracle.run_chain([ob.LZMH_DEC], junk, nbits, out_cap=1 << 20), (n_bytes, nbits)
    for period in range(1, 41):
        data = np.tile(rng.integers(0, 256, period, dtype=np.uint8), 700 // period + 1)[:700]
        rc, enc, nb = oracle.run_chain([ob.LZMH_ENC], data, data.size * 8)
        assert ctx.stage_run(ob.LZMH_DEC, enc, nb) == oracle.run_chain([ob.LZMH_DEC], enc, nb), period


def test_lzmh_batch_blocks(ctx, oracle):
    from data_compressor_b200 import synth
    data = synth.watts_u16(24, 32768 + 100, device="cuda").reshape(-1)       # 64 KiB blocks + a ragged tail
    payload, bit_len, byte_off, status, total, in_off = ctx.lzmh_encode(data, block_bytes=65536)
    assert not status.any().item()
    host = data.cpu().numpy()
    bl, bo, pay = bit_len.cpu().numpy(), byte_off.cpu().numpy(), payload[:total].cpu().numpy()
    for s in range(bl.size):
        blk = host[s * 65536:(s + 1) * 65536]
        rc, want, nb = oracle.run_chain([ob.LZMH_ENC], blk, blk.size * 8)
        assert nb == bl[s] and pay[int(bo[s]):int(bo[s]) + len(want)].tobytes() == want, s
    out, out_len, status = ctx.lzmh_decode(payload, total, bit_len, byte_off, in_off)
    # blocks that end in an un-ranked zero byte lose it (lzmh.c:571 quirk): compare against the oracle's decode
    for s in range(bl.size):
        o = int(bo[s])
        rc, want, nb = oracle.run_chain([ob.LZMH_DEC], pay[o:o + (int(bl[s]) + 7) // 8], int(bl[s]), out_cap=70000)
        assert int(out_len[s]) * 8 == nb
        assert out[int(in_off[s]):int(in_off[s]) + nb // 8].cpu().numpy().tobytes() == want


# ---- full size: BASELINE.json config 2 (4096 channels x 1 day), size-independent properties ------------------
@pytest.mark.parametrize("block_len", [1024, "bench"])
def test_full_size_round_trip_and_spot_checks(ctx, oracle, block_len):
    """The configuration bench.py runs ("bench" = the block length it derives from the device, 1168 on a 148-SM B200)
    and the 1024-sample blocks next to it: round trip of all samples, container invariants, oracle spot checks."""
    from data_compressor_b200 import synth
    nc, cl = 4096, 86400
    bl = ctx.two_wave_block_len(nc, cl) if block_len == "bench" else block_len
    per = -(-cl // bl)
    x = synth.redd_like(nc, cl, device="cuda")
    enc = ctx.dega_encode(x, block_len=bl)
    assert not enc.status.any().item()
    n = enc.n_streams
    assert n == nc * per
    if block_len == "bench":
        assert n <= 2 * ctx.coder_wave_streams() < n + nc        # two waves, as full as whole blocks per channel allow
    # container invariants: offsets are the exclusive scan of the byte counts
    nbytes = (enc.bit_len + 7) // 8
    assert torch.equal(enc.byte_off[1:], torch.cumsum(nbytes, 0)) and int(enc.byte_off[0]) == 0
    assert enc.total_bytes == int(enc.byte_off[-1])
    ratio = x.numel() * 4 / enc.total_bytes
    assert 2.0 < ratio < 4.0
    # lossless round trip of all 353,894,400 samples
    y, n_out, status = ctx.dega_decode(enc)
    assert not status.any().item()
    assert torch.equal(y.view(torch.int32), x.view(torch.int32))
    del y
    # spot-check random (channel, block) streams, and the ragged last block of a channel, bit for bit against the oracle
    rng = np.random.default_rng(0)
    for s in list(rng.integers(0, n, 24)) + [per - 1, n - 1]:
        c, b = divmod(int(s), per)
        xs = x[c, b * bl:(b + 1) * bl].cpu().numpy()
        rc, want, nb = oracle.dega_encode(xs)
        assert rc == 0 and nb == int(enc.bit_len[s]) and enc.stream_bytes(int(s)) == want


def test_whole_channel_streams_at_config_2_length(ctx, oracle):
    """block_len = 0: one stream per channel-day, the reference's own unit of work (one plain DCCLI run per channel,
    DCCLI/src/cli.c:430-466); 64 channels of the config-2 generator, every one compared with the oracle."""
    from data_compressor_b200 import synth
    x = synth.redd_like(64, 86400, device="cuda", seed=0xDC0000 + 640)
    enc = ctx.dega_encode(x, block_len=0)
    assert enc.n_streams == 64 and not enc.status.any().item()
    xs = x.cpu().numpy()
    for c in range(64):
        rc, want, nb = oracle.dega_encode(xs[c])
        assert rc == 0 and int(enc.bit_len[c]) == nb and enc.stream_bytes(c) == want, c
    y, n_out, status = ctx.dega_decode(enc)
    assert torch.equal(y.view(torch.int32), x.view(torch.int32)) and (n_out == 86400).all().item()


def test_lzmh_config_4_size_spot_checks(ctx, oracle):
    """BASELINE.json configs[3] at the bench's size: 16 384 blocks of 64 KiB (1 GiB); every block must come back, 32
    random blocks are compared with the oracle both ways."""
    from data_compressor_b200 import synth
    n_blocks, block = 16384, 65536
    data = synth.lzmh_blocks(n_blocks, block, device="cuda")
    payload, bit_len, byte_off, status, total, in_off = ctx.lzmh_encode(data, block_bytes=block)
    assert not status.any().item() and total == int(byte_off[-1])
    out, out_len, st2 = ctx.lzmh_decode(payload, total, bit_len, byte_off, in_off)
    assert not st2.any().item() and (out_len == block).all().item()       # no block ends in an un-ranked zero byte (synth.lzmh_blocks)
    assert torch.equal(out, data)
    rng = np.random.default_rng(4)
    for s in rng.integers(0, n_blocks, 32):
        s = int(s)
        blk = data[s * block:(s + 1) * block].cpu().numpy()
        rc, want, nb = oracle.run_chain([ob.LZMH_ENC], blk, blk.size * 8)
        o = int(byte_off[s])
        assert rc == 0 and nb == int(bit_len[s]) and payload[o:o + len(want)].cpu().numpy().tobytes() == want, s
        rc, back, nb2 = oracle.run_chain([ob.LZMH_DEC], want, nb, out_cap=block + 1024)
        assert rc == 0 and back == blk.tobytes()


# ---- regressions for the round-1 review ------------------------------------------------------------------------
def test_negative_normalization_factor_fails_like_the_reference(ctx, oracle):
    """normalization_factor is an unrestricted float option (DCLib/src/enc_dec.c:64-73): with a negative one the
    normalised values are negative and `encode diff` rejects them (diff.c:15-18).  The fused front end must not take
    its unsigned fast prefix for such samples."""
    rng = np.random.default_rng(9)
    x = profile(rng, (6, 300))
    x[2, :] = 0.0                   # zeros stay zero whatever the factor
    x[4, :] = 0.001                 # |x * nf| < 0.5: truncates to 0, no error
    for nf in (-100.0, -1.0, -0.5):
        enc = ctx.dega_encode(torch.from_numpy(x).cuda(), block_len=100, nf=nf)
        st = enc.status.cpu().numpy()
        for s, xs in enumerate(split_streams(x, 100)):
            rc, want, nb = oracle.dega_encode(xs, nf=nf)
            assert st[s] == rc, (nf, s)
            if rc == 0:
                assert enc.stream_bytes(s) == want and int(enc.bit_len[s]) == nb, (nf, s)
        for row in (0, 2, 4):
            rc, out, nb = ctx.stage_run(DEGA_ENC, x[row], None, 32, nf, True)
            want = oracle.run_chain(ob.DEGA_ENC, x[row], 300 * 32, 32, nf, 1)
            assert rc == want[0] and (rc != 0 or (rc, out, nb) == want), (nf, row)


def test_chain_buffers_survive_a_growing_staging_block(ctx, oracle):
    """dc_chain_run_host's ping-pong buffers and the staging block of the *_host calls are separate allocations:
    alternating the two with growing sizes must neither free nor reuse one for the other."""
    rng = np.random.default_rng(12)
    for n in (500, 40000, 2000, 300000, 1000):
        x = profile(rng, n)
        want = oracle.run_chain(ob.DEGA_ENC, x, n * 32, 32, 100.0, 1)
        got = ctx.chain_run(ob.DEGA_ENC, x, None, 32, 100.0, True)
        assert got[:3] == want, n
        big = profile(rng, (8, 4 * n))
        payload, bit_len, byte_off, status = ctx.dega_encode_host(big, block_len=0)
        assert not status.any()
        got = ctx.chain_run(ob.DEGA_ENC, x, None, 32, 100.0, True)
        assert got[:3] == want, n
        assert payload[int(byte_off[3]):int(byte_off[4])].tobytes() == oracle.dega_encode(big[3])[1]


def test_host_encode_many_chunks_with_unaligned_totals(ctx, oracle, monkeypatch):
    """Many small chunks whose payload sizes are not multiples of 16, pinned output buffers (so the copies back are
    really asynchronous): a chunk's payload must not land on the tail of the one before it on the device."""
    from data_compressor_b200 import synth
    monkeypatch.setenv("DC_HOST_CHUNK_STREAMS", "64")
    nc, cl, bl = 31, 1000, 40              # 775 streams -> 12 chunks of 2-3 channels
    x = synth.redd_like(nc, cl, device="cpu", seed=5).numpy()
    for rep in range(3):
        payload, bit_len, byte_off, status = ctx.dega_encode_host(x, block_len=bl)
        assert not status.any() and int(byte_off[-1]) == payload.size
        streams = split_streams(x, bl)
        for s, xs in enumerate(streams):
            want = oracle.dega_encode(xs)[1]
            o = int(byte_off[s])
            assert payload[o:o + len(want)].tobytes() == want, (rep, s)
        y, n_out, st2 = ctx.dega_decode_host(payload, bit_len, byte_off, nc, cl, bl)
        assert np.array_equal(y.view(np.uint32), x.view(np.uint32)) and not st2.any()


def test_decoders_reject_an_index_that_leaves_the_payload(ctx, oracle):
    """A truncated / corrupt stream index (byte range outside the payload) gives ERROR_INVALID_FORMAT for that stream,
    not a read past the allocation; the other streams still decode."""
    from data_compressor_b200.codec import Encoded
    rng = np.random.default_rng(21)
    x = profile(rng, (12, 400))
    enc = ctx.dega_encode(torch.from_numpy(x).cuda(), block_len=100)
    n = enc.n_streams
    bad_off = enc.byte_off.clone()
    bad_off[7] = enc.total_bytes + (1 << 40)          # far outside
    bad_len = enc.bit_len.clone()
    bad_len[20] = (enc.total_bytes + 5) * 8           # runs past the end
    bad_len[21] = -1                                   # 2^64 - 1 bits
    bad = Encoded(enc.payload, bad_len, bad_off, enc.status, enc.total_bytes, enc.n_channels, enc.channel_len, enc.block_len)
    y, n_out, status = ctx.dega_decode(bad)
    st = status.cpu().numpy()
    assert st[7] == -3 and st[20] == -3 and st[21] == -3
    good = np.ones(n, dtype=bool)
    good[[7, 20, 21]] = False
    assert not st[good].any()
    yy = y.cpu().numpy().reshape(-1, 100)
    assert np.array_equal(yy[good].view(np.uint32), x.reshape(-1, 100)[good].view(np.uint32))
    # the context is still usable (no sticky CUDA error)
    y2, _, st2 = ctx.dega_decode(enc)
    assert not st2.any().item() and torch.equal(y2.view(torch.int32).cpu(), torch.from_numpy(x).view(torch.int32))
    # LZMH: same rule
    data = torch.from_numpy(rng.integers(0, 7, 5000, dtype=np.uint8)).cuda()
    payload, bit_len, byte_off, status, total, in_off = ctx.lzmh_encode(data, block_bytes=1000)
    bo = byte_off.clone()
    bo[2] = total + 12345
    out, out_len, st3 = ctx.lzmh_decode(payload, total, bit_len, bo, in_off)
    assert st3.cpu().tolist() == [0, 0, -3, 0, 0] and int(out_len[2]) == 0
    assert out[:2000].cpu().numpy().tobytes() == data[:2000].cpu().numpy().tobytes()


def test_async_host_calls_overlap_an_encode_and_a_decode(ctx, oracle):
    """dc_dega_encode_host_async / dc_dega_decode_host_async / dc_wait: an encode and a decode in flight on ONE context
    (the job engine interleaves their chunks), jobs of the same kind queue up in order; results equal the blocking calls'."""
    from data_compressor_b200 import synth
    nc, cl, bl = 300, 3000, 250
    x = synth.redd_like(nc, cl, device="cpu", seed=31).numpy()
    x2 = synth.redd_like(nc, cl, device="cpu", seed=32).numpy()
    ref_enc = [np.array(a, copy=True) for a in ctx.dega_encode_host(x, block_len=bl)]
    ref2 = [np.array(a, copy=True) for a in ctx.dega_encode_host(x2, block_len=bl)]
    for rep in range(3):
        pe = ctx.dega_encode_host(x2, block_len=bl, wait=False, slot=1)
        pd = ctx.dega_decode_host(ref_enc[0], ref_enc[1], ref_enc[2], nc, cl, bl, wait=False)
        got2 = pe.wait()          # dc_wait: both jobs are finished after it
        y, n_out, st = pd.wait()
        assert np.array_equal(y.view(np.uint32), x.view(np.uint32)) and not st.any() and (n_out == bl).all()
        for a, b in zip(got2, ref2):
            assert np.array_equal(a, b)
    s = 1234
    c, b = divmod(s, cl // bl)
    want = oracle.dega_encode(x2[c, b * bl:(b + 1) * bl])[1]
    assert got2[0][int(got2[2][s]):int(got2[2][s]) + len(want)].tobytes() == want
    # two encodes queued behind each other on the same context: they run in order, each into its own buffers
    p1 = ctx.dega_encode_host(x, block_len=bl, wait=False, slot=0)
    p2 = ctx.dega_encode_host(x2, block_len=bl, wait=False, slot=1)
    g1, g2 = p1.wait(), p2.wait()
    for a, b in zip(g1, ref_enc):
        assert np.array_equal(a, b)
    for a, b in zip(g2, ref2):
        assert np.array_equal(a, b)
    assert ctx.lib.dc_wait(ctx.h) == 0                                # nothing in flight: a no-op
    # an asynchronous job that fails reports through dc_wait (payload capacity far too small)
    import ctypes as C
    par, lay = ctx._params(32, 100.0, True), ctx._layout(nc, cl, bl)
    n = ctx.n_streams(nc, cl, bl)
    small = np.zeros(64, dtype=np.uint8)
    blen, boff, stat, tot = np.zeros(n, np.uint64), np.zeros(n + 1, np.uint64), np.zeros(n, np.int32), C.c_uint64(0)
    assert ctx.lib.dc_dega_encode_host_async(ctx.h, x.ctypes.data, C.byref(lay), C.byref(par), small.ctypes.data, small.size, blen.ctypes.data,
                                             boff.ctypes.data, stat.ctypes.data, C.byref(tot)) == 0
    assert ctx.lib.dc_wait(ctx.h) == -6 and tot.value == ref_enc[0].size


# ---- aggregate (granularity reduction) ------------------------------------------------------------
def test_aggregate_matches_oracle(ctx, oracle):
    """dc_aggregate_device against the oracle (which is pinned to the reference's Aggregate() by
    tests/test_oracle_vs_ref.py): bit-exact float32 sums in the reference's order, per channel, short last group."""
    rng = np.random.default_rng(12)
    for nc, cl in ((1, 1), (3, 61), (7, 900), (2, 86400)):
        x = (rng.integers(0, 1 << 20, (nc, cl)).astype(np.float32) / np.float32(100.0)).astype(np.float32)
        x[rng.random((nc, cl)) < 0.05] *= np.float32(-1e-3)
        for k in (1, 2, 60, 900, 100000):
            got = ctx.aggregate(torch.from_numpy(x).cuda(), k).cpu().numpy()
            assert got.shape == (nc, (cl + k - 1) // k)
            for c in range(nc):
                want = oracle.aggregate(x[c], k)
                assert np.array_equal(got[c].view(np.uint32), want.view(np.uint32)), (nc, cl, k, c)


def test_aggregate_then_dega_on_coarser_granularity(ctx, oracle):
    """1 s -> 1 min on the device, then DEGA on the coarser series: same bytes as the oracle on the aggregated data."""
    from data_compressor_b200 import synth
    x = synth.redd_like(8, 7200, device="cuda")
    xm = ctx.aggregate(x, 60)
    # sums of sixty two-decimal values are not two-decimal floats any more; normalize with 1.0 like a Wh series
    enc = ctx.dega_encode(xm, block_len=0, nf=1.0)
    assert not enc.status.any().item()
    xs = xm.cpu().numpy()
    for c in range(8):
        rc, want, nb = oracle.dega_encode(xs[c], nf=1.0)
        assert rc == 0 and nb == int(enc.bit_len[c]) and enc.stream_bytes(c) == want


# ---- container -------------------------------------------------------------------------------------
def test_container_from_gpu_encode_round_trips(ctx, oracle, tmp_path):
    from data_compressor_b200 import container as ct
    rng = np.random.default_rng(31)
    x = profile(rng, (9, 3000))
    enc = ctx.dega_encode(torch.from_numpy(x).cuda(), block_len=1000)
    path = str(tmp_path / "gpu.dcb")
    h = ct.write_encoded(path, enc)
    assert (h.n_channels, h.channel_len, h.block_len, h.n_streams, h.payload_bytes) == (9, 3000, 1000, 27, enc.total_bytes)
    for s in (0, 13, 26):                                   # a stream cut out of the file is what the oracle / DCCLI writes
        data, nb = ct.stream_file(path, s)
        c, b = divmod(s, 3)
        rc, want, wb = oracle.dega_encode(x[c, b * 1000:(b + 1) * 1000])
        assert rc == 0 and (data, nb) == (want, wb)
    enc2, h2 = ct.load_encoded(path, "cuda")
    y, n_out, status = ctx.dega_decode(enc2, vs=h2.value_size_bits, nf=h2.normalization_factor, adaptive=bool(h2.adaptive))
    assert not status.any().item() and np.array_equal(y.cpu().numpy().view(np.uint32), x.view(np.uint32))


@pytest.mark.parametrize("nf", [100.0, 1.0, 3.0, 7.3, 0.01, 1000.0, 1e-5, 123456.789])
def test_denormalize_division_is_ieee_for_any_factor(ctx, oracle, nf):
    """The back end divides by the normalization factor with a reciprocal hoisted out of the sample loop (NfDivider in
    dega_device.cuh); the result has to be the correctly rounded float32 quotient the reference's divss gives, for
    small and large integers and for factors that are not exactly representable."""
    rng = np.random.default_rng(int(nf * 1000) % 1000 + 5)
    k = np.concatenate([np.arange(0, 300), rng.integers(0, 1 << 24, 3000), rng.integers(0, (1 << 31) - 300, 3000)]).astype(np.int64)
    k = np.abs(np.sort(k)[::-1] if nf == 3.0 else k)
    x = (k.astype(np.float32) / np.float32(nf)).astype(np.float32).reshape(4, -1)
    x = x[:, : x.shape[1] // 8 * 8]
    # keep only inputs the reference chain accepts at value size 32 (normalize.c:21-22, diff.c:17-18)
    v = x.astype(np.float32) * np.float32(nf) + np.float32(0.5)
    x[v >= np.float32(2147483000.0)] = np.float32(1.0) / np.float32(nf)
    enc = ctx.dega_encode(torch.from_numpy(x).cuda(), block_len=0, nf=nf)
    assert not enc.status.any().item()
    y, n_out, status = ctx.dega_decode(enc, nf=nf)
    assert not status.any().item()
    got = y.cpu().numpy()
    for c in range(x.shape[0]):
        rc, want = oracle.dega_decode(enc.stream_bytes(c), int(enc.bit_len[c]), nf=nf, out_cap=x.shape[1] * 4 + 64)
        assert rc == 0 and np.array_equal(got[c].view(np.uint32), want.view(np.uint32)), (nf, c)


# ---- csv -------------------------------------------------------------------------------------------
def _same_floats(a, b):
    return a.size == b.size and np.array_equal(a.view(np.uint32)[a == a], b.view(np.uint32)[b == b]) and np.array_equal(a != a, b != b)


def test_csv_decode_matches_oracle(ctx, oracle, golden):
    """dc_csv_decode_host against the oracle (pinned to the reference's ReadCSV by tests/test_oracle_vs_ref.py): quirks of
    the reference's character loop, fields strtof has to convert (exponents, hex, inf/nan, long digit strings), random
    decimals, and the reference's own test file re-printed with two decimals."""
    from test_oracle_vs_ref import CSV_CASES
    rng = np.random.default_rng(17)
    cases = list(CSV_CASES)
    for _ in range(20):
        n, d = int(rng.integers(1, 3000)), int(rng.integers(0, 10))
        rows = [",".join(f"{rng.integers(-10 ** 7, 10 ** 7) / 10 ** d:.{d}f}" for _ in range(int(rng.integers(1, 4)))) for _ in range(n)]
        cases.append("\n".join(rows) + ("\n" if rng.random() < 0.7 else ""))
    for t in cases:
        for col in (1, 2, 3):
            for sep in (",", ";"):
                assert _same_floats(ctx.csv_decode(t, col, sep), oracle.csv_decode(t, col, sep)), (t[:40], col, sep)
    text = "".join(f"{v:.2f}\n" for v in golden.testfile)              # DCCLI/testdata/input.txt has this form
    got = ctx.csv_decode(text)
    assert np.array_equal(got.view(np.uint32), golden.testfile.view(np.uint32))
    assert np.array_equal(got.view(np.uint32), oracle.csv_decode(text).view(np.uint32))


def test_csv_to_dega_pipeline(ctx, oracle, golden):
    """`decode csv # encode normalize # encode diff # encode seg # encode bac adaptive` with every stage on the GPU gives the
    bytes of the reference chain on the same text (169 327 bytes for the test file)."""
    text = "".join(f"{v:.2f}\n" for v in golden.testfile)
    x = ctx.csv_decode(text)
    enc = ctx.dega_encode(torch.from_numpy(x).cuda().view(1, -1), block_len=0)
    g = golden.get("testfile_stage4")
    assert int(enc.bit_len[0]) == g["out_bits"] and enc.stream_bytes(0) == g["out"].tobytes() and enc.total_bytes == 169327


def test_csv_encode_matches_oracle(ctx, oracle, golden):
    """dc_csv_encode_host against the oracle (pinned to the reference's WriteCSV in tests/test_oracle_vs_ref.py)."""
    from test_oracle_vs_ref import csv_encode_values
    x = csv_encode_values()
    big = np.array([3.4028235e38, -3.4028235e38, 1e38, 2.0 ** 127, 2.0 ** 100, 2.0 ** 64, 2.0 ** 63], dtype=np.float32)
    x = np.concatenate([x, big])                        # forces the capacity retry as well
    for d in range(7):
        for col, sep in ((1, ","), (4, ";")):
            assert ctx.csv_encode(x, col, sep, d) == oracle.csv_encode(x, col, sep, d), (d, col)
    assert ctx.csv_encode(np.zeros(0, np.float32)) == b""
    text = "".join(f"{v:.2f}\n" for v in golden.testfile).encode()
    assert ctx.csv_encode(golden.testfile) == text
    assert ctx.csv_encode(ctx.csv_decode(text)) == text                 # the ends of `make test`
    big_n = np.tile(golden.testfile, 40)                                # 4 M values: many CTAs, positions beyond 2^24
    assert ctx.csv_encode(big_n, 2, ",", 3) == oracle.csv_encode(big_n, 2, ",", 3)


def test_csv_decode_exponent_forms(ctx, oracle):
    """Decimal exponents are converted on the device when one correctly rounded operation does it; everything is compared
    with strtof (the oracle), whichever route a field takes."""
    rng = np.random.default_rng(9)
    rows = []
    for _ in range(20000):
        m = int(rng.integers(0, 10 ** int(rng.integers(1, 12))))
        frac = int(rng.integers(0, 8))
        e = int(rng.integers(-14, 15))
        body = f"{m / 10 ** frac:.{frac}f}" if rng.random() < 0.7 else str(m)
        rows.append(f"{'-' if rng.random() < 0.3 else ''}{body}{'eE'[int(rng.integers(0, 2))]}{'+' if e >= 0 and rng.random() < 0.5 else ''}{e}")
    rows += ["1e", "1e+", "1e-", "1.5e", "2E+x", "1e5x", "1p3", "12e00", "12e-00", "0e5", "0.0e-3", "1000e-3", "1e10", "1e11", "123456789e11",
             "1e38", "1e39", "1e-38", "1e-46", "9999999999999999999e0", "99999999999999999999e0", "16777217e0", "1.6777217e7", "1e99999999999"]
    text = "\n".join(rows) + "\n"
    got, want = ctx.csv_decode(text), oracle.csv_decode(text)
    assert got.size == len(rows) and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_csv_decode_many_strtof_fields(ctx, oracle):
    """More fields outside the fast grammar than the first pass keeps track of (65 536): the parse phase reruns with the
    exact list size and every one of them still gets strtof's value."""
    rng = np.random.default_rng(5)
    vals = rng.integers(1, 10 ** 6, 150000)
    text = "".join(f"0x{v:x}p-2\n" if i % 2 else f"{v / 100:.2f}\n" for i, v in enumerate(vals))      # hex floats: strtof's business
    got, want = ctx.csv_decode(text), oracle.csv_decode(text)
    assert got.size == vals.size and np.array_equal(got.view(np.uint32), want.view(np.uint32))


# ---- chains with the intermediate results on the device ------------------------------------------
@pytest.mark.parametrize("vs,adaptive", [(32, True), (24, False), (17, True), (64, True)])
def test_chain_run_matches_oracle_chain(ctx, oracle, golden, vs, adaptive):
    """dc_chain_run_host over the reference's command lines: the DEGA chain stage by stage and its inverse, partial chains,
    a chain that ends in lzmh, and the same chains cut after every stage (the intermediate results of a DCCLI run)."""
    x = golden.testfile[:20000]
    enc = [ob.NORMALIZE, ob.DIFF_ENC, ob.SEG_ENC, ob.BAC_ENC]
    dec = [ob.BAC_DEC, ob.SEG_DEC, ob.DIFF_DEC, ob.DENORMALIZE]
    for k in range(1, 5):
        want = oracle.run_chain(enc[:k], x, x.size * 32, vs=vs, adaptive=adaptive)
        got = ctx.chain_run(enc[:k], x, x.size * 32, vs=vs, adaptive=adaptive)
        assert got[:3] == want, (k, got[0], want[0])
    rc, payload, nb = oracle.run_chain(enc, x, x.size * 32, vs=vs, adaptive=adaptive)
    for k in range(1, 5):
        want = oracle.run_chain(dec[:k], payload, nb, vs=vs, adaptive=adaptive, out_cap=1 << 22)
        got = ctx.chain_run(dec[:k], payload, nb, vs=vs, adaptive=adaptive)
        assert got[:3] == want, (k, got[0], want[0])
    # eight stages: there and back in one call
    rc, back, nb2, _ = ctx.chain_run(enc + dec, x, x.size * 32, vs=vs, adaptive=adaptive)
    assert rc == 0 and nb2 == x.size * 32 and back == oracle.run_chain(enc + dec, x, x.size * 32, vs=vs, adaptive=adaptive, out_cap=1 << 22)[1]
    if vs == 32:
        assert np.array_equal(np.frombuffer(back, dtype=np.uint32), x.view(np.uint32))
        mixed = [ob.NORMALIZE, ob.DIFF_ENC, ob.LZMH_ENC]
        assert ctx.chain_run(mixed, x, x.size * 32)[:3] == oracle.run_chain(mixed, x, x.size * 32)
        assert ctx.chain_run(mixed + [ob.LZMH_DEC, ob.DIFF_DEC], x, x.size * 32)[:3] == \
            oracle.run_chain(mixed + [ob.LZMH_DEC, ob.DIFF_DEC], x, x.size * 32, out_cap=1 << 22)


def test_copy_stage_and_chains_that_contain_it(ctx, oracle):
    """`copy` (DCLib/src/copy.c:11-29) as a device stage: whole blocks pass unchanged, a short last block fails like the
    reference's checked read, and a chain that names it keeps its intermediate results on the device."""
    rng = np.random.default_rng(33)
    for bs, nbits in [(8, 800), (8, 803), (13, 13 * 50), (13, 700), (64, 6400), (1, 777), (24, 0), (0, 64)]:
        data = rng.integers(0, 256, (nbits + 7) // 8, dtype=np.uint8)
        if nbits % 8:
            data[-1] &= (0xFF00 >> (nbits % 8)) & 0xFF
        oracle.set_block_size(bs)
        want = oracle.run_chain([ob.COPY], data, nbits)
        got = ctx.stage_run(ob.COPY, data, nbits, blocksize=bs)
        assert got[0] == want[0], (bs, nbits)
        if want[0] == 0:
            assert got == want, (bs, nbits)
    x = profile(rng, 3000)
    oracle.set_block_size(32)
    chain = [ob.NORMALIZE, ob.COPY, ob.DIFF_ENC, ob.SEG_ENC, ob.COPY, ob.BAC_ENC]
    want = oracle.run_chain(chain, x, 3000 * 32, 32, 100.0, 1)
    oracle.set_block_size(8)
    # the SEG stream is not a whole number of 32-bit blocks in general: the second copy decides the outcome
    got = ctx.chain_run(chain, x, None, 32, 100.0, True, blocksize=32)
    assert got[0] == want[0] and (want[0] != 0 or got[:3] == want)
    want8 = oracle.run_chain(chain, x, 3000 * 32, 32, 100.0, 1)
    got8 = ctx.chain_run(chain, x, None, 32, 100.0, True, blocksize=1)
    oracle.set_block_size(1)
    want1 = oracle.run_chain(chain, x, 3000 * 32, 32, 100.0, 1)
    oracle.set_block_size(8)
    assert want1[0] == 0 and got8[:3] == want1 and (want8[0] != 0 or want8 == want1)


def test_chain_run_errors_and_edges(ctx, oracle):
    """The chain stops at the stage the reference would stop at, with its code; empty input; one stage; too many stages."""
    x = np.array([1.0, -2.5, 3.0], dtype=np.float32)            # a negative sample: diff refuses it (diff.c:17-18)
    chain = [ob.NORMALIZE, ob.DIFF_ENC, ob.SEG_ENC, ob.BAC_ENC]
    rc, out, nb, failed = ctx.chain_run(chain, x, 96)
    want = oracle.run_chain(chain, x, 96)
    assert rc == want[0] != 0 and failed == 1 and nb == 0
    big = np.array([3e9], dtype=np.float32)                      # out of range for valuesize 32: normalize refuses it
    rc, out, nb, failed = ctx.chain_run(chain, big, 32)
    assert rc == oracle.run_chain(chain, big, 32)[0] != 0 and failed == 0
    empty = np.zeros(0, dtype=np.float32)
    assert ctx.chain_run(chain, empty, 0)[:3] == oracle.run_chain(chain, empty, 0)
    assert ctx.chain_run([ob.SEG_ENC], np.arange(5, dtype=np.uint32), 160)[:3] == oracle.run_chain([ob.SEG_ENC], np.arange(5, dtype=np.uint32), 160)
    assert ctx.chain_run([ob.NORMALIZE] * 17, x, 96)[0] == -1    # ERROR_INVALID_VALUE: more than MAX_NUM_ENCODERS stages
    # a decoder chain whose first result is far larger than its input (the buffer between the stages has to grow)
    zeros = np.zeros(300000, dtype=np.float32)
    rc, payload, nb = oracle.run_chain(chain, zeros, zeros.size * 32, adaptive=True)
    dec = [ob.BAC_DEC, ob.SEG_DEC, ob.DIFF_DEC, ob.DENORMALIZE]
    got = ctx.chain_run(dec, payload, nb, adaptive=True)
    assert got[0] == 0 and got[2] == zeros.size * 32 and not np.frombuffer(got[1], dtype=np.uint32).any()
