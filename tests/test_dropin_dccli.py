"""Drop-in check at the reference's own front door: the unmodified DCCLI grammar, parameter parser, registry, csv
codec and DCIOLib buffers (compiled from /root/reference into oracle/_ref/), with the five hot-path codecs
replaced by data_compressor_b200/plugin/dc_cuda_plugin.c -> libdc_cuda.so.  Same command lines, same files."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPU = os.path.join(ROOT, "oracle", "_ref", "DCCLI")
GPU = os.path.join(ROOT, "oracle", "_ref", "DCCLI_cuda")


def run(binary, src, dst, chain):
    r = subprocess.run([binary, src, dst] + chain.split(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    return r.returncode


@pytest.fixture(scope="module")
def testfile(tmp_path_factory, golden):
    if not (os.path.exists(CPU) and os.path.exists(GPU)):
        pytest.skip("oracle/_ref/DCCLI[_cuda] did not travel to this box")
    d = tmp_path_factory.mktemp("dccli")
    p = d / "input.txt"
    # the reference's DCCLI/testdata/input.txt is one "%.2f" column; regenerate it from the committed float32 fixture
    p.write_text("".join(f"{v:.2f}\n" for v in golden.testfile))
    return d, str(p)


DEGA = "decode csv # encode normalize # encode diff # encode seg # encode bac adaptive"
UNDEGA = "decode bac adaptive # decode seg # decode diff # decode normalize # encode csv"


@pytest.mark.parametrize("chain", [
    DEGA,
    "decode csv # encode normalize",                                        # A-XDR
    "decode csv # encode normalize valuesize=24 # encode diff valuesize=24 # encode seg valuesize=24 # encode bac",
    "encode lzmh",
])
def test_same_command_line_same_bytes(testfile, chain):
    d, src = testfile
    a, b = str(d / "cpu.bin"), str(d / "gpu.bin")
    assert run(CPU, src, a, chain) == 0
    assert run(GPU, src, b, chain) == 0
    assert open(a, "rb").read() == open(b, "rb").read()


def test_make_test_through_the_gpu_dccli(testfile):
    """`make test` (DCCLI/build/gcc/Makefile:75-77): 10-stage DEGA round trip, then diff input output."""
    d, src = testfile
    out = str(d / "output.txt")
    assert run(GPU, src, out, DEGA + " # " + UNDEGA) == 0
    assert open(src, "rb").read() == open(out, "rb").read()


def test_errors_propagate_like_the_reference(testfile):
    d, src = testfile
    neg = str(d / "neg.txt")
    open(neg, "w").write("1.00\n-2.50\n3.00\n")
    a, b = str(d / "e_cpu.bin"), str(d / "e_gpu.bin")
    rc_cpu = run(CPU, neg, a, DEGA)
    rc_gpu = run(GPU, neg, b, DEGA)
    assert rc_cpu == rc_gpu != 0          # ERROR_LIBRARY_CALL as exit status (cli.c:447-453)


@pytest.mark.parametrize("options", ["", "column=2", "column=2 separator_char=;", "column=3 separator_char=;"])
def test_csv_reader_behind_the_unchanged_registry_row(testfile, options):
    """`decode csv ...` of DCCLI_cuda is the plugin's ReadCSV_CUDA (oracle/Makefile binds the "csv" row's reader to it);
    ragged rows, both separators, exponent / hex / inf fields that take the strtof route, no newline at the end."""
    d, _ = testfile
    src = str(d / "table.csv")
    rows = ["1.5;2.25;3", "4;5.75", "6;7;8", "1,2,3", "-0.25,1e3,0x1p4", ",inf,", "", "12345678.9,0.000001,99", "7,8"]
    open(src, "w").write("\n".join(rows))
    a, b = str(d / "c_cpu.bin"), str(d / "c_gpu.bin")
    chain = ("decode csv " + options).strip()
    assert run(CPU, src, a, chain) == 0
    assert run(GPU, src, b, chain) == 0
    assert open(a, "rb").read() == open(b, "rb").read() and os.path.getsize(a) > 0


def run_env(binary, src, dst, chain, **env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([binary, src, dst] + chain.split(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, env=e)
    return r.returncode, r.stderr.decode(errors="replace")


@pytest.mark.parametrize("chain,pieces", [
    (DEGA + " # " + UNDEGA, 3),                                   # csv | eight GPU stages as one chain | csv
    (DEGA, 2),
    ("encode normalize # encode copy # encode diff # encode copy blocksize=32 # encode seg # encode bac adaptive", 1),
    ("encode lzmh # decode lzmh # encode copy blocksize=16", 1),
    ("encode copy blocksize=24", None),                           # a single stage: nothing to fuse, the reference's own main runs it
])
def test_chain_fuser_in_front_of_the_stage_loop(testfile, golden, chain, pieces):
    """dc_fused_main.c (linked in front of the unmodified cli.c) runs every run of consecutive GPU stages of a command
    line as ONE device-resident chain; the result is the file the stage-by-stage path (DC_CUDA_NO_FUSE) and the CPU
    DCCLI write, and the report says how many pieces the command line was cut into."""
    d, src = testfile
    if not chain.startswith("decode csv"):
        src = str(d / "input.f32")
        golden.testfile[:6000].tofile(src)      # raw float32 in (24 000 bytes: whole 16-, 24- and 32-bit blocks)
    a, b, c = str(d / "f_cpu.bin"), str(d / "f_fused.bin"), str(d / "f_staged.bin")
    assert run(CPU, src, a, chain) == 0
    rc, err = run_env(GPU, src, b, chain, DC_CUDA_FUSE_REPORT="1")
    assert rc == 0, err
    if pieces is not None:
        assert f"{len(chain.split('#'))} stages in {pieces} pieces" in err
    else:
        assert "dc_fused_main" not in err
    rc, err = run_env(GPU, src, c, chain, DC_CUDA_NO_FUSE="1")
    assert rc == 0, err
    assert open(a, "rb").read() == open(b, "rb").read() == open(c, "rb").read()


def test_chain_fuser_reports_the_failing_stage_like_the_reference(testfile):
    d, _ = testfile
    src = str(d / "odd.bin")
    open(src, "wb").write(bytes(range(7)))       # 56 bits: not a whole number of 32-bit blocks
    chain = "encode copy # encode copy blocksize=32 # encode copy"
    rc_cpu = run(CPU, src, str(d / "o_cpu.bin"), chain)
    rc_gpu, err = run_env(GPU, src, str(d / "o_gpu.bin"), chain)
    assert rc_cpu == rc_gpu != 0
    assert "encoder/decoder 2 of 3" in err
