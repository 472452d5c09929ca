"""The oracle (oracle/dc_oracle.c) against the golden vectors produced by the real reference."""
import hashlib

import numpy as np
import pytest


def test_golden_inventory(golden):
    assert len(golden.names) >= 80
    assert golden.testfile.size == 100000
    # sha256 of `decode csv` output quoted in SURVEY.md section 4
    assert hashlib.sha256(golden.testfile.tobytes()).hexdigest().startswith("a865ae0245bf6a46")


def test_oracle_matches_every_golden_vector(oracle, golden):
    for name in golden.names:
        g = golden.get(name)
        rc, out, nb = oracle.run_chain(g["chain"], g["input"], g["in_bits"], g["vs"], g["nf"], g["adaptive"],
                                       out_cap=max(1 << 16, g["input"].size * 40))
        assert rc == g["rc"], name
        assert nb == g["out_bits"], name
        assert hashlib.sha256(out).hexdigest() == g["sha256"], name
        if g["out"].size:
            assert out == g["out"].tobytes(), name


def test_survey_known_answers(oracle, golden):
    """sizes and hashes of the reference's own test file."""
    want = {"testfile_stage1": (3200000, "385d3fa1a6c98f74"), "testfile_stage2": (3200000, "0aba5da8105b3f62"),
            "testfile_stage3": (1566518, "3ae049e58af2040f"), "testfile_stage4": (1354613, "721f7ed9e8757f43"),
            "testfile_bac_static": (2482828, "f0425eb9209a2f4a"), "lzmh_full_text": (270896 * 8 - 2, "c493269c5c7b98da")}
    for name, (bits, sha) in want.items():
        g = golden.get(name)
        assert g["out_bits"] == bits and g["sha256"].startswith(sha), name
    assert golden.get("small_stage3")["out"].tobytes() == bytes.fromhex("019003200ca0")
    assert golden.get("small_stage4")["out"].tobytes() == bytes.fromhex("f9c008475e84")
    assert golden.get("empty_dega")["out"].tobytes() == bytes.fromhex("20")
    assert golden.get("lzmh_A")["out"].tobytes() == bytes.fromhex("1040")


def test_make_test_round_trip(oracle, golden):
    """The reference's only test (`make test`, DCCLI/build/gcc/Makefile:75-77): DEGA round trip of the test file."""
    x = golden.testfile
    rc, pay, nb = oracle.dega_encode(x, adaptive=1)
    assert rc == 0 and (nb + 7) // 8 == 169327
    rc, y = oracle.dega_decode(pay, nb, adaptive=1)
    assert rc == 0 and np.array_equal(y.view(np.uint32), x.view(np.uint32))


@pytest.mark.parametrize("vs", [17, 24, 32, 64])
def test_seg_output_independent_of_valuesize(oracle, golden, vs):
    x = golden.testfile[:5000]
    a = oracle.dega_encode(x, vs=32)
    b = oracle.dega_encode(x, vs=vs)
    assert a == b
