"""Randomised parity of the oracle against the REAL reference (oracle/_ref, compiled from /root/reference).
Skipped where the reference build is absent."""
import numpy as np
import pytest

from oracle import binding as ob


def rand_profile(rng, n, scale=300.0, neg=False):
    k = np.cumsum(rng.integers(-300, 301, n)) + int(scale * 100)
    k = np.abs(k) if not neg else k
    return (k.astype(np.float32) / np.float32(100.0)).astype(np.float32)


@pytest.mark.parametrize("seed", range(6))
def test_dega_chain_random(oracle, ref, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 5000))
    x = rand_profile(rng, n)
    vs = int(rng.choice([20, 24, 32, 40, 64]))
    ad = int(rng.integers(0, 2))
    for k in range(1, 5):
        chain = ob.DEGA_ENC[:k]
        assert oracle.run_chain(chain, x, n * 32, vs, 100.0, ad) == ref.run_chain(chain, x, n * 32, vs, 100.0, ad)
    rc, pay, nb = oracle.dega_encode(x, vs, 100.0, ad)
    for k in range(1, 5):
        chain = ob.DEGA_DEC[:k]
        assert oracle.run_chain(chain, pay, nb, vs, 100.0, ad) == ref.run_chain(chain, pay, nb, vs, 100.0, ad)


@pytest.mark.parametrize("seed", range(4))
def test_error_codes_random(oracle, ref, seed):
    rng = np.random.default_rng(100 + seed)
    x = rand_profile(rng, 200, neg=True)
    x[int(rng.integers(0, 200))] = np.float32(rng.choice([-5.0, 1e9, -1e12, 400000.0]))
    for vs in (16, 24, 32):
        for k in range(1, 5):
            a = oracle.run_chain(ob.DEGA_ENC[:k], x, None, vs, 100.0, 1)
            b = ref.run_chain(ob.DEGA_ENC[:k], x, None, vs, 100.0, 1)
            assert a[0] == b[0] and (a[0] != 0 or a == b)


@pytest.mark.parametrize("seed", range(4))
def test_bac_random_bits(oracle, ref, seed):
    rng = np.random.default_rng(200 + seed)
    n = int(rng.integers(0, 60000))
    p1 = float(rng.choice([0.01, 0.2, 0.5, 0.9]))
    bits = np.packbits((rng.random(n) < p1).astype(np.uint8))
    for ad in (0, 1):
        a = oracle.run_chain([ob.BAC_ENC], bits, n, adaptive=ad)
        assert a == ref.run_chain([ob.BAC_ENC], bits, n, adaptive=ad)
        assert oracle.run_chain([ob.BAC_DEC], a[1], a[2], adaptive=ad) == ref.run_chain([ob.BAC_DEC], a[1], a[2], adaptive=ad)
        # garbage in: both must agree on the error / output as well
        junk = rng.integers(0, 256, 64, dtype=np.uint8)
        assert oracle.run_chain([ob.BAC_DEC], junk, adaptive=ad)[0] == ref.run_chain([ob.BAC_DEC], junk, adaptive=ad)[0]


@pytest.mark.parametrize("seed", range(4))
def test_seg_decode_garbage(oracle, ref, seed):
    rng = np.random.default_rng(300 + seed)
    junk = rng.integers(0, 256, 200, dtype=np.uint8)
    nbits = int(rng.integers(1, 1600))
    for vs in (8, 17, 32, 64):
        a = oracle.run_chain([ob.SEG_DEC], junk, nbits, vs)
        b = ref.run_chain([ob.SEG_DEC], junk, nbits, vs)
        assert a[0] == b[0] and (a[0] != 0 or a == b)


@pytest.mark.parametrize("seed", range(6))
def test_lzmh_random(oracle, ref, seed):
    rng = np.random.default_rng(400 + seed)
    n = int(rng.choice([0, 1, 2, 3, 127, 128, 129, 274, 275, 402, 403, 404, 1000, 5000]))
    kind = seed % 3
    if kind == 0:
        data = rng.integers(0, 4, n, dtype=np.uint8)                       # highly repetitive
    elif kind == 1:
        data = np.frombuffer(("".join(f"{v:.2f}\n" for v in rng.random(n // 5 + 1) * 300))[:n].encode(), dtype=np.uint8)
    else:
        data = rng.integers(0, 256, n, dtype=np.uint8)
    a = oracle.run_chain([ob.LZMH_ENC], data, data.size * 8)
    b = ref.run_chain([ob.LZMH_ENC], data, data.size * 8)
    assert a == b
    assert oracle.run_chain([ob.LZMH_DEC], a[1], a[2]) == ref.run_chain([ob.LZMH_DEC], a[1], a[2])


def test_aggregate_oracle_vs_reference(oracle, ref):
    """`encode aggregate num_values=k` (aggregate.c:9-27): sequential float32 sums, short last group, empty input."""
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 3, 59, 60, 61, 900, 4097):
        for k in (1, 2, 3, 15, 60, 900):
            x = (rng.integers(0, 1 << 20, n).astype(np.float32) / np.float32(100.0)).astype(np.float32)
            x[rng.random(n) < 0.05] *= np.float32(-1e-3)      # mixed signs and magnitudes: the order of the adds matters
            a, b = oracle.aggregate(x, k), ref.aggregate(x, k)
            assert a.size == (n + k - 1) // k and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (n, k)


CSV_CASES = [
    "1.00\n2.00\n1.50\n", "1.5", "", "\n", "\n\n", "1,2,3\n4,5\n6,7,8", "1,2,", "1,2,\n", "a,b\n,,\n\n5", "12.5\r\n13\r\n",
    "-0\n+3.25\n 7\n1e3\n0x10\ninf\nnan\n-.5\n5.\n.\n-\n123456789\n0.000012345\n1234.500000\n16777217\n3.4028236e38\n",
    "  \t9.75x\n1.2.3\n--4\n+-4\n1e\n1e+\n0x\n0xg\nINFINITY\nNaN(12)\n00012.50\n-000.0100\n99999999999999999999\n0.1234567891\n",
    "3;4;5\n6;7\n", "7\n8",
    "1e\n1e+\n1.5e\n2E+x\n1p3\n12e00\n0e5\n1000e-3\n1e10\n1e11\n123456789e11\n1e38\n1e39\n1e-38\n1e-46\n16777217e0\n1.6777217e7\n-2.5E-3\n",
]


def _same_floats(a, b):
    return a.size == b.size and np.array_equal(a.view(np.uint32)[a == a], b.view(np.uint32)[b == b]) and np.array_equal(a != a, b != b)


def test_csv_decode_oracle_vs_reference(oracle, ref):
    """`decode csv` (csv.c:13-44): field splitting, column selection, the end-of-input rules and strtof's grammar."""
    rng = np.random.default_rng(17)
    cases = list(CSV_CASES)
    for _ in range(30):     # random decimals with 0..9 fraction digits, sometimes three columns, sometimes no last newline
        n, d = int(rng.integers(1, 200)), int(rng.integers(0, 10))
        rows = []
        for _ in range(n):
            vals = [f"{rng.integers(-10 ** 7, 10 ** 7) / 10 ** d:.{d}f}" for _ in range(int(rng.integers(1, 4)))]
            rows.append(",".join(vals))
        cases.append("\n".join(rows) + ("\n" if rng.random() < 0.7 else ""))
    for t in cases:
        for col in (1, 2, 3):
            for sep in (",", ";"):
                assert _same_floats(oracle.csv_decode(t, col, sep), ref.csv_decode(t, col, sep)), (t[:40], col, sep)


def csv_encode_values():
    """Values for `encode csv`: ties of the decimal rounding, signed zeros, denormals, non-finite values, integers beyond
    2^64, meter-like readings and random bit patterns.  |x| < 1e30: the reference's text buffer (csv.c:11) holds 48
    characters, which larger values with six decimals overflow."""
    rng = np.random.default_rng(23)
    special = np.array([0.0, -0.0, 0.5, 1.5, 2.5, 0.125, 0.375, -0.005, 0.005, 1e-7, 3.4e29, -3.4e29, 1e-45, -1e-45, np.inf, -np.inf, np.nan,
                        16777216.0, 1e10, 1.8446744e19, 3.6893488e19, 123456.789, 0.045, 0.055, 2.0 ** -24, 1 - 2.0 ** -24, 9.995, 99.995,
                        0.9999999, 0.95, 0.25, 0.75, 8388608.5, 4194304.25], dtype=np.float32)
    neg_nan = np.array([0xFFC00000, 0x7F800001, 0xFF800001], dtype=np.uint32).view(np.float32)
    rnd = rng.integers(0, 2 ** 32, 20000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    rnd = rnd[~(np.abs(rnd) >= 1e30)]
    meter = (rng.integers(0, 10 ** 7, 20000) / 100).astype(np.float32)
    halves = (rng.integers(0, 2 ** 20, 5000) / 2.0 ** rng.integers(1, 12, 5000)).astype(np.float32)      # exact binary fractions: real ties
    return np.concatenate([special, neg_nan, rnd, meter, halves])


def test_csv_encode_oracle_vs_reference(oracle, ref):
    """`encode csv` (csv.c:46-65): printf's rounding, the column prefix and the separator."""
    x = csv_encode_values()
    for d in range(7):
        for col, sep in ((1, ","), (3, ";")):
            assert oracle.csv_encode(x, col, sep, d) == ref.csv_encode(x, col, sep, d), (d, col)
    assert oracle.csv_encode(np.zeros(0, np.float32)) == ref.csv_encode(np.zeros(0, np.float32)) == b""


def test_csv_round_trip_of_the_test_file(oracle, ref, golden):
    """decode csv # encode csv gives the test file's text back (what `make test` diffs at the end)."""
    text = "".join(f"{v:.2f}\n" for v in golden.testfile)
    assert ref.csv_encode(ref.csv_decode(text)) == text.encode() == oracle.csv_encode(oracle.csv_decode(text))


def test_copy_stage_matches_the_reference(oracle, ref):
    """`encode copy blocksize=..` (DCLib/src/copy.c:11-29): whole blocks pass, a short last block is a failed read after
    the whole blocks before it were written."""
    rng = np.random.default_rng(77)
    try:
        for bs in (1, 5, 8, 13, 32, 64, 100, 1000):
            for nbits in (0, bs, 3 * bs, 3 * bs + 1, 10 * bs - 1, 997):
                data = rng.integers(0, 256, (nbits + 7) // 8, dtype=np.uint8)
                if nbits % 8:
                    data[-1] &= (0xFF00 >> (nbits % 8)) & 0xFF
                oracle.set_block_size(bs)
                ref.set_block_size(bs)
                assert oracle.run_chain([ob.COPY], data, nbits) == ref.run_chain([ob.COPY], data, nbits), (bs, nbits)
    finally:
        oracle.set_block_size(8)
        ref.set_block_size(8)
