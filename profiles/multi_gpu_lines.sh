#!/bin/bash
# usage: multi_gpu_lines.sh N [sweep]  -- runs the multi-GPU bench lines (and the configs[4] sweep) on N GPUs of the box
N=$1
mkdir -p gpurun_out/mg
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/mg/dega_$N.json 2> gpurun_out/mg/dega_$N.err; echo "dega rc=$?"; tail -c 300 gpurun_out/mg/dega_$N.err
$TR bench.py --gpus $N --workload axdr --steps 10 --no-cpu > gpurun_out/mg/axdr32_$N.json 2> gpurun_out/mg/axdr32_$N.err; echo "axdr32 rc=$?"
$TR bench.py --gpus $N --workload axdr --valuesize 24 --steps 10 --no-cpu > gpurun_out/mg/axdr24_$N.json 2> gpurun_out/mg/axdr24_$N.err; echo "axdr24 rc=$?"
$TR bench.py --gpus $N --workload lzmh --steps 5 --no-cpu > gpurun_out/mg/lzmh_$N.json 2> gpurun_out/mg/lzmh_$N.err; echo "lzmh rc=$?"
if [ "$2" = "sweep" ]; then
  $TR bench.py --gpus $N --workload sweep --steps 3 --cpu-seconds 3 > gpurun_out/mg/sweep_$N.jsonl 2> gpurun_out/mg/sweep_$N.err; echo "sweep rc=$?"; wc -l gpurun_out/mg/sweep_$N.jsonl
fi
python - <<PY
import json
for w in ("dega","axdr32","axdr24","lzmh"):
    try:
        d=json.loads(open(f"gpurun_out/mg/{w}_$N.json").read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(w, "n_gpus", d["n_gpus"], "value %.1f" % d["value"], d["unit"], "ms %.3f" % d["ms_per_step"], "ok", d.get("round_trip_ok"), "e2e", e.get("value"), "dec", d.get("decode_gbs"))
    except Exception as ex:
        print(w, "FAILED", ex)
PY
