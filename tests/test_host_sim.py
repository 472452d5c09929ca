"""Logic tests of the kernels' per-thread bodies, compiled for the host by tests/host_sim (no GPU needed).

This is the same source that nvcc compiles for sm_100a (data_compressor_b200/csrc/*.cuh); the harness is test-only
and is not a fallback path.  The GPU-side parity tests proper are in test_gpu_parity.py (-m gpu)."""
import numpy as np
import pytest

from conftest import split_streams
from oracle import binding as ob

# enum dc_stage == enum dco_stage_id for the ten plain stages; 10/11 are the fused DEGA chains
DEGA_ENC, DEGA_DEC = 10, 11


def test_division_table_is_exact(sim):
    assert sim.lib.sim_check_div_table(65521) == 0


def test_every_golden_vector_stage_by_stage(sim, golden):
    """Runs each golden chain one stage at a time through the device bodies, like DCCLI's stage loop."""
    for name in golden.names:
        g = golden.get(name)
        if any(s in (ob.LZMH_ENC, ob.LZMH_DEC) for s in g["chain"]):
            continue
        data, nbits, rc = g["input"].tobytes(), g["in_bits"], 0
        for st in g["chain"]:
            rc, data, nbits = sim.stage_run(st, data, nbits, g["vs"], g["nf"], g["adaptive"])
            if rc != 0:
                break
        assert rc == g["rc"], name
        if rc == 0:
            assert nbits == g["out_bits"], name
            if g["out"].size:
                assert data == g["out"].tobytes(), name


def test_fused_dega_matches_golden(sim, golden):
    g = golden.get("testfile_stage4")
    rc, out, nb = sim.stage_run(DEGA_ENC, golden.testfile, None, 32, 100.0, 1)
    assert rc == 0 and nb == g["out_bits"] and out == g["out"].tobytes()
    rc, dec, nb2 = sim.stage_run(DEGA_DEC, out, nb, 32, 100.0, 1, cap=golden.testfile.size * 4 + 64)
    assert rc == 0 and dec == golden.testfile.tobytes()


@pytest.mark.parametrize("seed", range(8))
def test_random_streams_vs_oracle(sim, oracle, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(0, 3000))
    k = np.abs(np.cumsum(rng.integers(-int(rng.choice([3, 300, 30000])), 301, n)) + 20000)
    if seed % 2:   # appliance-like jumps: code words of 33 bits and more (|delta| >= 2^15) between the short ones
        jumps = rng.random(n) < 0.02
        k = np.abs(k + np.cumsum(np.where(jumps, rng.integers(-3_000_000, 3_000_001, n), 0))) % (1 << 20)   # fits every value size drawn below
    x = (k.astype(np.float32) / np.float32(100.0)).astype(np.float32)
    vs = int(rng.choice([22, 24, 32, 33, 48, 64]))
    ad = int(rng.integers(0, 2))
    want = oracle.run_chain(ob.DEGA_ENC, x, n * 32, vs, 100.0, ad)
    got = sim.stage_run(DEGA_ENC, x, n * 32, vs, 100.0, ad)
    assert got == want
    back = sim.stage_run(DEGA_DEC, got[1], got[2], vs, 100.0, ad, cap=n * 4 + 64)
    assert back == oracle.run_chain(ob.DEGA_DEC, got[1], got[2], vs, 100.0, ad)
    if k.size == 0 or k.max() < (1 << 21):     # float32 round trip is only exact for moderate magnitudes
        assert back[1] == x.tobytes()


@pytest.mark.parametrize("nf", [-100.0, -1.0, -0.25, 0.0])
def test_fused_front_end_with_non_positive_factors(sim, oracle, nf):
    """normalization_factor is unrestricted (DCLib/src/enc_dec.c:64-73).  A negative one makes the normalised values
    negative, which `encode diff` rejects (diff.c:15-18); tiny products truncate to 0 and pass.  The fused front end's
    unsigned fast prefix must hand such samples to the general path."""
    rng = np.random.default_rng(7)
    cases = [np.array([1.0, 2.0, 1.5], dtype=np.float32), np.zeros(40, dtype=np.float32), np.full(40, 0.001, dtype=np.float32),
             (np.abs(rng.integers(0, 5000, 200)).astype(np.float32) / np.float32(100.0)).astype(np.float32),
             np.concatenate([np.zeros(9, dtype=np.float32), np.array([0.004, 3.0], dtype=np.float32)])]
    for x in cases:
        want = oracle.run_chain(ob.DEGA_ENC, x, x.size * 32, 32, nf, 1)
        got = sim.stage_run(DEGA_ENC, x, x.size * 32, 32, nf, 1)
        assert got[0] == want[0], (nf, x[:4])       # a failing chain leaves no defined output
        if want[0] == 0:
            assert got == want, (nf, x[:4])


@pytest.mark.parametrize("seed", range(6))
def test_bac_random_bits_vs_oracle(sim, oracle, seed):
    rng = np.random.default_rng(50 + seed)
    n = int(rng.choice([0, 1, 2, 31, 32, 33, 1000, 20000, 70000]))
    p1 = float(rng.choice([0.001, 0.1, 0.5, 0.97]))
    bits = np.packbits((rng.random(n) < p1).astype(np.uint8))
    for ad in (0, 1):
        want = oracle.run_chain([ob.BAC_ENC], bits, n, adaptive=ad)
        got = sim.stage_run(ob.BAC_ENC, bits, n, adaptive=ad)
        assert got == want
        assert sim.stage_run(ob.BAC_DEC, got[1], got[2], adaptive=ad) == oracle.run_chain([ob.BAC_DEC], got[1], got[2], adaptive=ad)
        # the decoder reading the byte-padded file sees 8 * nbytes bits
        assert sim.stage_run(ob.BAC_DEC, got[1], len(got[1]) * 8, adaptive=ad) == \
            oracle.run_chain([ob.BAC_DEC], got[1], len(got[1]) * 8, adaptive=ad)


def test_bac_long_runs_force_halvings_and_pending_bits(sim, oracle):
    # 200k symbols: > 20 halvings; alternating blocks produce long E3 (pending) runs
    rng = np.random.default_rng(7)
    bits = np.concatenate([np.zeros(60000, np.uint8), (rng.random(80000) < 0.5).astype(np.uint8), np.ones(60000, np.uint8)])
    packed = np.packbits(bits)
    redo0 = [sim.lib.sim_bac_redo_count(i) for i in (0, 1)]
    want = oracle.run_chain([ob.BAC_ENC], packed, bits.size, adaptive=1)
    got = sim.stage_run(ob.BAC_ENC, packed, bits.size, adaptive=1)
    assert got == want
    assert sim.stage_run(ob.BAC_DEC, got[1], got[2], adaptive=1)[1] == packed.tobytes()
    # the first ones after 60000 zeros cost ~13 bits each: the unrolled 16-symbol groups run out of room (31 / 32 bits)
    # and are redone symbol by symbol, in the encoder and in the decoder
    assert all(sim.lib.sim_bac_redo_count(i) > redo0[i] for i in (0, 1))


def test_bac_long_pending_runs(sim, oracle):
    """Static model: a '1' always takes the middle third, so runs of ones pile up E3 steps (pending > 16 takes the
    generic emitter in bac.cuh); then a '0' or the end of the stream resolves them."""
    for n_ones in (5, 11, 12, 20, 40, 100, 1000):
        for tail in ([], [0], [0, 1, 1, 0]):
            bits = np.array([1] * n_ones + tail, dtype=np.uint8)
            packed = np.packbits(bits)
            for ad in (0, 1):
                want = oracle.run_chain([ob.BAC_ENC], packed, bits.size, adaptive=ad)
                got = sim.stage_run(ob.BAC_ENC, packed, bits.size, adaptive=ad)
                assert got == want, (n_ones, tail, ad)
                back = sim.stage_run(ob.BAC_DEC, got[1], got[2], adaptive=ad)
                assert back[0] == 0 and back[2] == bits.size and back[1] == packed.tobytes()


@pytest.mark.parametrize("seed", range(4))
def test_garbage_decoders_vs_oracle(sim, oracle, seed):
    rng = np.random.default_rng(90 + seed)
    junk = rng.integers(0, 256, 300, dtype=np.uint8)
    nbits = int(rng.integers(1, 2400))
    for vs in (8, 17, 32, 64):
        a = sim.stage_run(ob.SEG_DEC, junk, nbits, vs)
        b = oracle.run_chain([ob.SEG_DEC], junk, nbits, vs)
        assert a[0] == b[0] and (a[0] != 0 or a == b)
    for ad in (0, 1):
        a = sim.stage_run(ob.BAC_DEC, junk, nbits, adaptive=ad)
        b = oracle.run_chain([ob.BAC_DEC], junk, nbits, adaptive=ad)
        assert a[0] == b[0] and (a[0] != 0 or a == b)


@pytest.mark.parametrize("block_len", [0, 1, 7, 64, 1000])
def test_batch_layout_and_dense_payload(sim, oracle, block_len):
    """(channel, block) streams, ragged last block, dense byte-aligned payload, unaligned decode reads."""
    rng = np.random.default_rng(block_len)
    nc, cl = 5, 333
    k = np.abs(np.cumsum(rng.integers(-200, 201, (nc, cl)), axis=1) + 30000)
    x = (k.astype(np.float32) / np.float32(100.0)).astype(np.float32)
    rc, payload, bit_len, byte_off, status = sim.dega_encode_batch(x, block_len)
    assert rc == 0 and not status.any()
    streams = split_streams(x, block_len)
    assert len(streams) == bit_len.size
    for s, xs in enumerate(streams):
        orc, obytes, obits = oracle.dega_encode(xs)
        assert orc == 0 and obits == bit_len[s]
        o = int(byte_off[s])
        assert payload[o:o + len(obytes)].tobytes() == obytes
        assert int(byte_off[s + 1]) - o == len(obytes)
    rc, y, n_out, status = sim.dega_decode_batch(payload, bit_len, byte_off, nc, cl, block_len)
    assert rc == 0 and not status.any()
    assert np.array_equal(y.view(np.uint32), x.view(np.uint32))
    assert [int(v) for v in n_out] == [len(s) for s in streams]


def test_batch_failed_stream_is_isolated(sim, oracle):
    x = np.full((3, 50), 12.5, dtype=np.float32)
    x[1, 20] = -3.0           # negative sample: diff reads it unsigned and rejects
    rc, payload, bit_len, byte_off, status = sim.dega_encode_batch(x, 0)
    assert list(status) == [0, -1, 0] and bit_len[1] == 0
    assert payload[int(byte_off[2]):].tobytes() == oracle.dega_encode(x[2])[1]


@pytest.mark.parametrize("adaptive", [0, 1])
def test_bac_carry_through_stored_words(sim, oracle, adaptive):
    """Pending runs of 30..300 bits resolved upwards and downwards: the encoder's carry has to ripple through its
    accumulator and through words already stored (CarryWriter::ripple in bac.cuh)."""
    from conftest import straddling_bits
    rng = np.random.default_rng(3 + adaptive)
    ff = zz = 0
    for n in (1, 2, 28, 29, 30, 31, 32, 33, 34, 40, 61, 62, 63, 64, 65, 66, 90, 127, 128, 129, 200, 300):
        for tail in ([], [0], [1], [0, 0, 0], [1, 1, 1], list(rng.integers(0, 2, 40))):
            for prefix, resolve in (((), "up"), ((), "down"), (tuple(rng.integers(0, 2, int(rng.integers(1, 70)))), None)):
                bits = straddling_bits(n, tail, adaptive, prefix, resolve)
                packed = np.packbits(bits)
                want = oracle.run_chain([ob.BAC_ENC], packed, bits.size, adaptive=adaptive)
                got = sim.stage_run(ob.BAC_ENC, packed, bits.size, adaptive=adaptive)
                assert got == want, (n, tail, prefix, resolve)
                back = sim.stage_run(ob.BAC_DEC, got[1], got[2], adaptive=adaptive)
                assert back[0] == 0 and back[2] == bits.size and back[1] == packed.tobytes()
                ff += b"\xff" * 6 in want[1]     # a run resolved downwards: 0111...
                zz += b"\x00" * 6 in want[1]     # a run resolved by a carry: 1000...
    assert adaptive == 0 or (ff > 20 and zz > 20)   # the static model cannot straddle for long (EOF owns the lowest third)


def test_csv_print_matches_oracle(sim, oracle):
    """The per-value body of `encode csv` (csv_device.cuh) against the oracle's printf, all decimal counts."""
    from test_oracle_vs_ref import csv_encode_values
    x = csv_encode_values()
    big = np.array([3.4028235e38, -3.4028235e38, 1e38, 2.0 ** 127, 2.0 ** 100, 2.0 ** 64, 2.0 ** 63, 1.8446743e19], dtype=np.float32)
    x = np.concatenate([x, big])                        # the oracle's buffer, unlike the reference's, holds these
    for d in range(7):
        for col, sep in ((1, ","), (4, ";")):
            assert sim.csv_encode(x, col, sep, d) == oracle.csv_encode(x, col, sep, d), (d, col)


def test_csv_field_conversion_matches_strtof(sim, oracle):
    """csv_to_float (csv_device.cuh) against strtof for every field it does not hand over, and the hand-over itself for
    the forms that need it."""
    rng = np.random.default_rng(31)
    fields = ["0", "-0", "+0.0", "1", "1.5", "-2.25", " 7", "\t8.5", "00012.50", "-000.0100", "5.", ".5", "-.5", ".", "-", "", "abc", "1.2.3", "--4", "+-4",
              "9.75x", "123456789", "1234567890123456789", "12345678901234567890", "0.000012345", "1234.500000", "16777216", "16777217", "16777219",
              "0.1", "0.12345678", "0.1234567891", "99999999", "3.4028235e38", "3.4028236e38", "1e", "1e+", "1e-", "1.5e", "2E+x", "1e5x", "1p3", "12e00",
              "12e-00", "0e5", "0.0e-3", "1000e-3", "1e10", "1e11", "123456789e11", "1e38", "1e39", "1e-38", "1e-46", "9999999999999999999e0",
              "16777217e0", "1.6777217e7", "1e99999999999", "-2.5E-3", "0x10", "0x1p4", "0x", "0xg", "inf", "-inf", "INFINITY", "nan", "NaN(12)", "1e-10", "1e-11"]
    for _ in range(30000):
        s = str(int("".join(str(d) for d in rng.integers(0, 10, int(rng.integers(1, 21))))))      # up to 20 digits
        frac = int(rng.integers(0, 12))
        body = (s[:-frac] or "0") + "." + s[-frac:].rjust(frac, "0") if frac else s
        if rng.random() < 0.4:
            body += "eE"[int(rng.integers(0, 2))] + ("+" if rng.random() < 0.3 else "") + str(int(rng.integers(-14, 15)))
        fields.append(("-" if rng.random() < 0.3 else "") + body)
    n_fast = 0
    for f in fields:
        v, hard = sim.csv_field(f)
        want = oracle.csv_decode(f + "\n")
        assert want.size == 1
        if not hard:
            n_fast += 1
            assert v.view(np.uint32) == want.view(np.uint32)[0], (f, v, want[0])
    assert all(sim.csv_field(f)[1] for f in ("0x10", "inf", "nan", "12345678901234567890", "1e39", "1e-46", "0.1234567891"))
    assert not any(sim.csv_field(f)[1] for f in ("1.5", "123456789", "1e3", "-2.5E-3", "1e", "1p3", "", "abc", "1000e-3"))
    assert n_fast > 5000


def test_one_register_interval_update_agrees_with_the_top_aligned_one(sim):
    """BacEncoder::narrow_w against BacEncoder::narrow_fast on 20 M random (window, increment, range) triples: same number
    of doublings, same new range and window, same outgoing bits (carry included)."""
    assert sim.lib.sim_check_narrow_w(20_000_000, 1) == 0


def test_dense_reader_matches_a_plain_bit_string(sim):
    """The arithmetic decoder's reader (bitio.cuh: DenseReader): every byte phase, stream ends anywhere in a word, bits
    past the end read as zero, the stream that follows in the payload is never seen."""
    import ctypes as C
    rng = np.random.default_rng(77)
    for trial in range(300):
        buf = rng.integers(0, 256, 64, dtype=np.uint8)          # neighbours are random bits, not zeros
        byte_off = int(rng.integers(0, 8))
        nbits = int(rng.integers(0, 8 * (48 - byte_off)))
        allbits = np.unpackbits(buf)[8 * byte_off: 8 * byte_off + nbits]
        steps = rng.integers(0, 33, 24, dtype=np.uint32)
        if trial % 3 == 0:
            steps[:] = 32
        peeks = np.zeros(steps.size, dtype=np.uint32)
        pos = sim.lib.sim_dense_reader(buf.ctypes.data, buf.size, byte_off, nbits, steps.ctypes.data, steps.size, peeks.ctypes.data)
        at = 0
        for i, st in enumerate(steps):
            window = np.zeros(32, dtype=np.uint8)
            have = allbits[at:at + 32]
            window[:have.size] = have
            want = int(np.packbits(window).view(">u4")[0])
            assert int(peeks[i]) == want, (trial, i, byte_off, nbits, at)
            at += int(st)
        assert pos == at


def test_carry_writer_is_big_integer_addition(sim):
    """The encoder's output stage (bac.cuh: CarryWriter): appending k bits with a carry bit above them is
    X = (X << k) + v on an arbitrarily long integer -- carries through bits still pending and through stored words.
    The sequences are what a coder can produce: the number written so far is the leading bits of a fixed target T, or one
    below them (the low end of a shrinking interval creeps up on T), so the bits in front of any position change at most
    once, by +1 -- the writer's precondition."""
    import ctypes as C
    rng = np.random.default_rng(5)
    for trial in range(400):
        n = int(rng.integers(1, 60))
        ks = rng.integers(0, 32, n).astype(np.uint32)
        vs = np.zeros(n, dtype=np.uint32)
        total = int(ks.sum())
        if trial % 2:
            T = int.from_bytes(rng.bytes(total // 8 + 1), "big") >> (8 - total % 8) if total else 0
        else:   # long runs of zeros and ones in the target: a lag of one below it is a long run of ones, the carry ripples far
            T = 0
            for _ in range(total):
                T = (T << 1) | (1 if (T & 1 and rng.random() < 0.95) or (not T & 1 and rng.random() < 0.05) else 0)
        T &= (1 << total) - 1 if total else 0
        X, L, lag = 0, 0, 0
        for i in range(n):
            k = int(ks[i])
            lead = T >> (total - L - k)                      # the first L + k bits of the target
            new_lag = int(rng.integers(0, 2))
            if lead - new_lag < (X << k):                    # the number never goes down (and never below zero)
                new_lag = 0
            v = lead - new_lag - (X << k)
            assert 0 <= v < (2 << k)
            vs[i] = v
            X, L, lag = lead - new_lag, L + k, new_lag
        out = np.zeros(((L + 31) // 32 + 2) * 4, dtype=np.uint8)
        ov = C.c_int(0)
        bits = sim.lib.sim_carry_writer(vs.ctypes.data, ks.ctypes.data, n, out.ctypes.data, out.size, C.byref(ov))
        assert bits == L and ov.value == 0
        want = np.zeros(out.size * 8, dtype=np.uint8)
        if L:
            want[:L] = [int(c) for c in format(X, "0%db" % L)]
        assert np.array_equal(np.unpackbits(out)[:((L + 31) // 32) * 32], want[:((L + 31) // 32) * 32]), trial
