"""Container file: format round trip, position-based shard writes, and the drop-in property --
a stream cut out of the container is the file the reference's DCCLI reads.  CPU only: payloads come from the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import split_streams
from oracle import binding as ob


def _profiles(rng, nc, cl):
    k = np.abs(np.cumsum(rng.integers(-300, 301, (nc, cl)), axis=1) + 30000)
    return (k.astype(np.float32) / np.float32(100.0)).astype(np.float32)


def _oracle_encode(oracle, x, block_len):
    streams = split_streams(x, block_len)
    pays, bits = [], []
    for xs in streams:
        rc, pay, nb = oracle.dega_encode(xs)
        assert rc == 0
        pays.append(pay)
        bits.append(nb)
    return streams, pays, np.array(bits, dtype=np.uint64)


def test_container_round_trip_and_shard_writes(oracle, tmp_path):
    from data_compressor_b200 import container as ct
    rng = np.random.default_rng(1)
    x = _profiles(rng, 6, 500)
    streams, pays, bits = _oracle_encode(oracle, x, 128)
    payload = np.frombuffer(b"".join(pays), dtype=np.uint8)
    path = str(tmp_path / "a.dcb")
    h = ct.make_header(ct.CHAIN_DEGA, len(streams), payload.size, 6, 500, 128, 32, 100.0, True)
    ct.create(path, h)
    # three "ranks" write their slices out of order
    off = ct.byte_offsets(bits)
    cuts = [0, 5, 11, len(streams)]
    for r in (2, 0, 1):
        a, b = cuts[r], cuts[r + 1]
        ct.write_shard(path, h, a, bits[a:b], int(off[a]), payload[int(off[a]):int(off[b])])
    assert os.path.getsize(path) == h.payload_offset + payload.size and h.payload_offset % 64 == 0
    g = ct.read_header(path)
    assert (g.chain, g.value_size_bits, g.adaptive, g.n_channels, g.channel_len, g.block_len, g.n_streams, g.payload_bytes) == \
        (1, 32, 1, 6, 500, 128, len(streams), payload.size) and g.normalization_factor == 100.0
    assert np.array_equal(ct.read_index(path), bits) and np.array_equal(ct.read_payload(path), payload)
    assert np.array_equal(ct.read_index(path, g, 3, 4), bits[3:7])
    for s in (0, 7, len(streams) - 1):
        data, nb = ct.stream_file(path, s)
        assert data == pays[s] and nb == int(bits[s])
        rc, y = oracle.dega_decode(data, nb)
        assert rc == 0 and np.array_equal(y.view(np.uint32), streams[s].view(np.uint32))


def test_container_rejects_foreign_and_truncated_files(tmp_path):
    from data_compressor_b200 import DcError, container as ct
    p = str(tmp_path / "junk")
    open(p, "wb").write(b"\x00" * 500)
    with pytest.raises(DcError):
        ct.read_header(p)
    h = ct.make_header(ct.CHAIN_DEGA, 4, 100, 1, 40, 10)
    ct.create(p, h)
    os.truncate(p, h.payload_offset + 50)
    with pytest.raises(DcError):
        ct.read_header(p)
    with pytest.raises(DcError):
        ct.read_header(str(tmp_path / "missing"))
    with pytest.raises(DcError):
        ct.write_shard(p, h, 3, np.zeros(2, np.uint64), 0, np.zeros(1, np.uint8))     # index slice past the end


def test_container_header_fields_cannot_wrap(tmp_path):
    """Corrupt headers whose sums would wrap 64 bits are rejected, as are slices given with wrapping arguments."""
    import ctypes as C
    from data_compressor_b200 import DcError, _lib, container as ct
    lib = _lib.load()
    p = str(tmp_path / "c.dcb")
    h = ct.make_header(ct.CHAIN_DEGA, 4, 100, 1, 40, 10)
    ct.create(p, h)
    raw = bytearray(open(p, "rb").read())
    for field, value in (("n_streams", (1 << 61) + 1), ("payload_bytes", (1 << 64) - 8), ("index_offset", (1 << 64) - 64)):
        g = _lib.ContainerHeader.from_buffer_copy(bytes(raw[:128]))
        setattr(g, field, value)
        q = str(tmp_path / f"bad_{field}")
        open(q, "wb").write(bytes(g) + bytes(raw[128:]))
        with pytest.raises(DcError):
            ct.read_header(q)
        assert lib.dc_container_read_index(os.fsencode(q), C.byref(g), 0, 1, np.zeros(1, np.uint64).ctypes.data) == -1
    buf = np.zeros(4, np.uint64)
    assert lib.dc_container_read_index(os.fsencode(p), C.byref(h), (1 << 64) - 1, 2, buf.ctypes.data) == -1
    assert lib.dc_container_read_payload(os.fsencode(p), C.byref(h), (1 << 64) - 4, 8, buf.ctypes.data) == -1
    assert lib.dc_container_write_shard(os.fsencode(p), C.byref(h), 2, (1 << 64) - 1, buf.ctypes.data, 0, None, 0) == -1


def test_load_encoded_checks_the_index_against_the_payload(oracle, tmp_path):
    """An index whose byte counts do not add up to the header's payload size is refused before any decoder sees it."""
    torch = pytest.importorskip("torch")
    from data_compressor_b200 import DcError, container as ct
    rng = np.random.default_rng(3)
    x = _profiles(rng, 2, 300)
    streams, pays, bits = _oracle_encode(oracle, x, 100)
    payload = np.frombuffer(b"".join(pays), dtype=np.uint8)
    path = str(tmp_path / "d.dcb")
    h = ct.make_header(ct.CHAIN_DEGA, len(streams), payload.size, 2, 300, 100)
    ct.create(path, h)
    ct.write_shard(path, h, 0, bits, 0, payload)
    enc, g = ct.load_encoded(path, device="cpu")
    assert enc.total_bytes == payload.size and np.array_equal(enc.bit_len.numpy().astype(np.uint64), bits)
    bad = bits.copy()
    bad[2] += 4096
    ct.write_shard(path, h, 0, bad, 0, payload)
    with pytest.raises(DcError):
        ct.load_encoded(path, device="cpu")


def test_stream_cut_out_of_the_container_decodes_with_the_reference_cli(oracle, tmp_path):
    """dd one stream out of the container, run the reference's own DCCLI decode chain on it: the samples come back."""
    cli = ob.REF_CLI
    if not os.path.exists(cli):
        pytest.skip("oracle/_ref/DCCLI not available")
    from data_compressor_b200 import container as ct
    rng = np.random.default_rng(2)
    x = _profiles(rng, 3, 700)
    streams, pays, bits = _oracle_encode(oracle, x, 256)
    path = str(tmp_path / "b.dcb")
    h = ct.make_header(ct.CHAIN_DEGA, len(streams), sum(len(p) for p in pays), 3, 700, 256)
    ct.create(path, h)
    ct.write_shard(path, h, 0, bits, 0, np.frombuffer(b"".join(pays), dtype=np.uint8))
    off = ct.byte_offsets(bits)
    for s in (0, 4, len(streams) - 1):
        cut, out = str(tmp_path / f"s{s}.bin"), str(tmp_path / f"s{s}.f32")
        subprocess.run(["dd", f"if={path}", f"of={cut}", "bs=1", f"skip={h.payload_offset + int(off[s])}",
                        f"count={(int(bits[s]) + 7) // 8}", "status=none"], check=True)
        r = subprocess.run([cli, cut, out, "decode", "bac", "adaptive", "#", "decode", "seg", "#", "decode", "diff", "#", "decode", "normalize"],
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr
        y = np.fromfile(out, dtype=np.float32)
        assert np.array_equal(y.view(np.uint32), streams[s].view(np.uint32))
