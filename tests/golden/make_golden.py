"""Regenerates tests/golden/*.npz from the REAL reference (oracle/_ref/libdcref.so, built from /root/reference by
oracle/Makefile).  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Every vector stores the input, the chain, the parameters, the reference's return code and its exact output bits.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TESTFILE = "/root/reference/DataCompressor/DCCLI/testdata/input.txt"


def main():
    ob.build(ref=True)
    ref = ob.Ref()
    vectors = []

    def add(name, data, nbits, chain, vs=32, nf=100.0, adaptive=1, keep_out=True):
        src = np.frombuffer(bytes(data), dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else \
            np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        rc, out, nb = ref.run_chain(chain, src, nbits, vs, nf, adaptive, out_cap=max(1 << 16, src.size * 40))
        vectors.append(dict(name=name, input=src.copy(), in_bits=nbits, chain=np.array(chain, dtype=np.int32), vs=vs, nf=nf,
                            adaptive=adaptive, rc=rc, out_bits=nb,
                            out=np.frombuffer(out, dtype=np.uint8).copy() if keep_out else np.zeros(0, np.uint8),
                            sha256=hashlib.sha256(out).hexdigest()))

    # ---- the reference's own test file (DCCLI/testdata/input.txt), csv -> float32 exactly like `decode csv` (strtof)
    txt = open(TESTFILE).read().split()
    x = np.array([np.float32(t) for t in txt], dtype=np.float32)
    np.save(os.path.join(HERE, "testfile_f32.npy"), x)
    full = [ob.NORMALIZE, ob.DIFF_ENC, ob.SEG_ENC, ob.BAC_ENC]
    for k in range(1, 5):
        add(f"testfile_stage{k}", x, x.size * 32, full[:k], keep_out=(k == 4))
    add("testfile_bac_static", x, x.size * 32, full, adaptive=0, keep_out=False)
    dega = vectors[3]
    add("testfile_decode", dega["out"], dega["out_bits"], ob.DEGA_DEC, keep_out=False)

    # ---- small known answers
    small = np.array([1.0, 2.0, 1.5], dtype=np.float32)
    for k in range(1, 5):
        add(f"small_stage{k}", small, 96, full[:k])
    add("empty_dega", b"", 0, full)
    add("empty_bac_static", b"", 0, [ob.BAC_ENC], adaptive=0)

    # ---- value sizes, first 3000 samples
    xs = x[:3000]
    for vs in (17, 24, 32, 48, 64):
        add(f"vs{vs}_dega", xs, xs.size * 32, full, vs=vs)
        add(f"vs{vs}_norm", xs, xs.size * 32, [ob.NORMALIZE], vs=vs)
        v = vectors[-2]
        add(f"vs{vs}_decode", v["out"], v["out_bits"], ob.DEGA_DEC, vs=vs)
    # vs = 16: 327.67 normalises to 32767.5 > 32767 -> rejected
    add("vs16_reject", np.array([1.0, 327.67], dtype=np.float32), 64, [ob.NORMALIZE], vs=16)
    add("vs16_ok", np.array([1.0, 327.66, 0.0, 12.34], dtype=np.float32), 128, full, vs=16)
    # negative sample: normalize accepts, diff reads it unsigned and rejects
    add("negative_norm", np.array([1.0, -2.5, 3.0], dtype=np.float32), 96, [ob.NORMALIZE])
    add("negative_chain", np.array([1.0, -2.5, 3.0], dtype=np.float32), 96, full)
    add("nf10", xs[:500], 500 * 32, full, nf=10.0)
    add("nf1000", xs[:500], 500 * 32, full, nf=1000.0)
    # short read: 5 bytes into normalize
    add("short_float", bytes([0, 0, 128, 63, 1]), 40, [ob.NORMALIZE])

    # ---- random bit strings through bac (both models), including a long skewed one that triggers halving
    rng = np.random.default_rng(1234)
    for i, (n, p1) in enumerate([(1, 0.5), (7, 0.5), (64, 0.5), (1000, 0.3), (5000, 0.05), (40000, 0.02), (70000, 0.5)]):
        bits = (rng.random(n) < p1).astype(np.uint8)
        packed = np.packbits(bits)
        for ad in (0, 1):
            add(f"bac_rand{i}_a{ad}", packed, n, [ob.BAC_ENC], adaptive=ad)
            v = vectors[-1]
            add(f"bac_rand{i}_a{ad}_dec", v["out"], v["out_bits"], [ob.BAC_DEC], adaptive=ad)
    # decoding from the byte-padded FILE (8 * nbytes bits) must give the same symbols
    v = [q for q in vectors if q["name"] == "bac_rand3_a1"][0]
    add("bac_padded_file_dec", v["out"], v["out"].size * 8, [ob.BAC_DEC], adaptive=1)
    # truncated stream -> ERROR_INVALID_FORMAT once more than 14 bits past the end are needed
    add("bac_truncated_dec", v["out"][: v["out"].size // 2], (v["out"].size // 2) * 8, [ob.BAC_DEC], adaptive=1)
    add("bac_empty_dec", b"", 0, [ob.BAC_DEC], adaptive=1)

    # ---- seg decoder edge cases
    add("seg_padding", bytes([0b01010110, 0b10000000]), 16, [ob.SEG_DEC])          # trailing zeros are padding
    add("seg_prefix_then_eof", bytes([0b10100000]), 4, [ob.SEG_DEC])               # '1','01',then '0' + EOF
    add("seg_prefix_one_eof", bytes([0b01000000]), 2, [ob.SEG_DEC])                # '01' then EOF: silent stop
    add("seg_short_postfix", bytes([0b00100000]), 4, [ob.SEG_DEC])                 # '001' + 1 of 2 postfix bits
    add("seg_too_long", bytes(8), 64, [ob.SEG_DEC], vs=32)                         # 33 zeros -> invalid format

    # ---- LZMH
    raw = open(TESTFILE, "rb").read()
    add("lzmh_text20k", raw[:20000], 160000, [ob.LZMH_ENC])
    v = vectors[-1]
    add("lzmh_text20k_dec", v["out"], v["out_bits"], [ob.LZMH_DEC])
    add("lzmh_full_text", raw, len(raw) * 8, [ob.LZMH_ENC], keep_out=False)
    add("lzmh_A", b"A", 8, [ob.LZMH_ENC])
    add("lzmh_A0", b"A\0", 16, [ob.LZMH_ENC])
    add("lzmh_A0_dec", bytes([0x10, 0x40, 0x00]), 20, [ob.LZMH_DEC])               # trailing zero literal is dropped
    add("lzmh_403", raw[:403], 403 * 8, [ob.LZMH_ENC])                             # the 403-byte quirk: empty output
    add("lzmh_404", raw[:404], 404 * 8, [ob.LZMH_ENC])
    add("lzmh_zeros", bytes(1000), 8000, [ob.LZMH_ENC])
    w = (np.clip(np.round(x[:4000]), 0, 65535)).astype("<u2").tobytes()
    add("lzmh_u16", w, len(w) * 8, [ob.LZMH_ENC])
    v = vectors[-1]
    add("lzmh_u16_dec", v["out"], v["out_bits"], [ob.LZMH_DEC])
    rb = rng.integers(0, 256, 3000, dtype=np.uint8).tobytes()
    add("lzmh_random", rb, len(rb) * 8, [ob.LZMH_ENC])
    add("lzmh_random_dec", vectors[-1]["out"], vectors[-1]["out_bits"], [ob.LZMH_DEC])

    pack = {}
    names = []
    for v in vectors:
        n = v["name"]
        names.append(n)
        for key in ("input", "chain", "out"):
            pack[f"{n}/{key}"] = v[key]
        pack[f"{n}/meta"] = np.array([v["in_bits"], v["vs"], v["adaptive"], v["rc"], v["out_bits"]], dtype=np.int64)
        pack[f"{n}/nf"] = np.array([v["nf"]], dtype=np.float32)
        pack[f"{n}/sha256"] = np.frombuffer(v["sha256"].encode(), dtype=np.uint8)
    # the big test-file input is stored once as testfile_f32.npy
    for n in names:
        if n.startswith("testfile_stage") or n == "testfile_bac_static":
            pack[f"{n}/input"] = np.zeros(0, np.uint8)
    pack["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **pack)
    for v in vectors:
        print(f"{v['name']:28s} rc={v['rc']:4d} in_bits={v['in_bits']:9d} out_bits={v['out_bits']:9d} {v['sha256'][:16]}")


if __name__ == "__main__":
    main()
