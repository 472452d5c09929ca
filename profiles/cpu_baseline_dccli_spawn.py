"""The CPU baseline in the form BASELINE.md / SURVEY.md 8(d) states it: the reference's release build, raw float32 stream
files on tmpfs, ONE DCCLI PROCESS PER STREAM (encode chain, then decode chain), `xargs -P $(nproc)` over all streams.
Reports wall time, aggregate GB/s of raw samples for encode + decode, per-core MB/s and the sum of DCCLI's own
"Total time elapsed" prints (cli.c:338-352: clock() around every stage).  bench.py's cpu_baseline times the same codecs
through in-process calls (one worker per core) instead; this script calibrates that substitute.

    python profiles/cpu_baseline_dccli_spawn.py [--streams N] [--block-len L]"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

DCCLI = os.path.join(ROOT, "oracle", "_ref", "DCCLI")
ENC = "encode normalize # encode diff # encode seg # encode bac adaptive"
DEC = "decode bac adaptive # decode seg # decode diff # decode normalize"
SI = {"n": 1e-9, "u": 1e-6, "m": 1e-3, "": 1.0, "k": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--block-len", type=int, default=1168)
    args = ap.parse_args()
    cores = os.cpu_count() or 1
    n = args.streams or cores * 250
    from data_compressor_b200 import synth
    per_channel = -(-86400 // args.block_len)
    x = synth.redd_like(-(-n // per_channel), 86400, device="cpu").numpy()
    tmp = "/dev/shm/dc_spawn" if os.path.isdir("/dev/shm") else "/tmp/dc_spawn"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    files, raw = [], 0
    for s in range(n):
        c, b = divmod(s, per_channel)
        blk = x[c, b * args.block_len:(b + 1) * args.block_len]
        f = os.path.join(tmp, f"s{s:07d}.f32")
        blk.tofile(f)
        files.append(f)
        raw += blk.size * 4
    q = lambda chain: chain.replace("#", "'#'")     # a bare # starts a comment in sh
    script = f'{DCCLI} "$0" "$0.enc" {q(ENC)} > "$0.log" && {DCCLI} "$0.enc" "$0.dec" {q(DEC)} >> "$0.log"'
    t0 = time.perf_counter()
    r = subprocess.run(["xargs", "-P", str(cores), "-n", "1", "sh", "-c", script], input="\n".join(files).encode(), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-500:]
    cpu_s, payload = 0.0, 0
    for f in files:
        for m in re.finditer(r"Total time elapsed: ([0-9.]+) ?([numk]?)s", open(f + ".log").read()):
            cpu_s += float(m.group(1)) * SI[m.group(2)]
        payload += os.path.getsize(f + ".enc")
    for f in files[:: max(1, n // 50)]:      # lossless round trip, spot checked
        assert open(f, "rb").read() == open(f + ".dec", "rb").read(), f
    print(json.dumps({
        "what": "reference release build, one DCCLI process per stream (encode chain then decode chain), xargs -P nproc, files on tmpfs",
        "streams": n, "block_len": args.block_len, "raw_bytes": raw, "payload_bytes": payload, "compression_ratio": raw / payload,
        "cores": cores, "wall_s": wall, "value": raw / wall / 1e9, "unit": "GB/s of raw samples, encode + decode", "per_core_mb_s": raw / wall / 1e6 / cores,
        "dccli_stage_seconds_summed": cpu_s, "gbs_from_stage_seconds": raw / (cpu_s / cores) / 1e9 if cpu_s else None,
        "processes": 2 * n}))
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
