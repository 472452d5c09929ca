#!/bin/bash
# One B200: every bench line of the round (written to gpurun_out/r2/), the reference arm and the DCCLI-spawn CPU baseline.
mkdir -p gpurun_out/r2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_default.json 2> gpurun_out/r2/bench_default.err; echo "default rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/bench_reference.json 2> gpurun_out/r2/bench_reference.err; echo "reference rc=$?"
python bench.py --workload axdr --steps 10 > gpurun_out/r2/bench_axdr32.json 2> gpurun_out/r2/bench_axdr32.err; echo "axdr32 rc=$?"
python bench.py --workload axdr --valuesize 24 --steps 10 > gpurun_out/r2/bench_axdr24.json 2> gpurun_out/r2/bench_axdr24.err; echo "axdr24 rc=$?"
python bench.py --workload lzmh --steps 5 > gpurun_out/r2/bench_lzmh.json 2> gpurun_out/r2/bench_lzmh.err; echo "lzmh rc=$?"
python bench.py --workload csv --steps 5 > gpurun_out/r2/bench_csv.json 2> gpurun_out/r2/bench_csv.err; echo "csv rc=$?"
python bench.py --workload sweep --steps 3 --cpu-seconds 3 > gpurun_out/r2/bench_sweep.jsonl 2> gpurun_out/r2/bench_sweep.err; echo "sweep rc=$?"
python profiles/cpu_baseline_dccli_spawn.py > gpurun_out/r2/cpu_dccli_spawn.json 2> gpurun_out/r2/cpu_dccli_spawn.err; echo "spawn rc=$?"
tail -c 400 gpurun_out/r2/*.err
