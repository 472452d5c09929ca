"""bench.py's reference arm runs on host cores only, so its JSON line -- the keys the driver reads -- can be checked
without a GPU; the product arm must refuse to run without one (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line(oracle):
    env = dict(os.environ, DC_BENCH_REF_SECONDS="0.3")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "0.2"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    line = json.loads(r.stdout.decode().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "DEGA encode+decode GB/s of raw samples" and line["unit"] == "GB/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "4096 channels x 86400 samples" in line["config"]["workload"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       timeout=600, cwd=ROOT)
    assert r.returncode != 0 and b"no CPU fallback" in r.stderr + r.stdout
