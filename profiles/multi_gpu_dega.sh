#!/bin/bash
# usage: multi_gpu_dega.sh N  -- the default (config 2) bench line on N GPUs of the box
N=$1
mkdir -p gpurun_out/mg
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/mg/dega_$N.json 2> gpurun_out/mg/dega_$N.err; echo "dega rc=$?"; tail -c 300 gpurun_out/mg/dega_$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/mg/dega_$N.json").read().strip().splitlines()[-1])
e=d.get("e2e") or {}
print("n_gpus", d["n_gpus"], "value %.1f" % d["value"], "ms %.3f" % d["ms_per_step"], "ok", d.get("round_trip_ok"), "e2e", e.get("value"), "floor", e.get("pcie_floor_ms"), "kernel_ms", d.get("kernel_ms"))
PY
