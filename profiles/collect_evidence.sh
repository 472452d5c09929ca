#!/bin/bash
# One B200: the whole GPU test suite, then the ncu evidence of the round (launch list of the bench command, --set full of the DEGA kernels).
# The launch list is filtered to the library's kernels (the first 600 launches of the process are otherwise torch's input generator).
# Every ncu pass runs only after the same command has exited 0 without ncu; nothing printed under ncu is a bench value.
mkdir -p gpurun_out/ev
python -m pytest tests -m gpu -q 2>&1 | tail -3
CMD="python bench.py --steps 4 --warmup 3 --no-cpu --e2e-stream-steps 4"
$CMD > gpurun_out/ev/plain.json 2> gpurun_out/ev/plain.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(front_|back_|bac_|scan_|gather_|publish_|validate_)" -c 600 --csv --log-file gpurun_out/ev/launches.csv $CMD > gpurun_out/ev/ncu_bench.log 2>&1; echo "launch list rc=$?"
python profiles/prof_step.py 4096 1168 1 > gpurun_out/ev/step_plain.log 2>&1; echo "step rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"^(bac_encode|bac_decode|front_dega|back_dega|gather)" -c 5 -o gpurun_out/ev/dega python profiles/prof_step.py 4096 1168 1 > gpurun_out/ev/ncu_step.log 2>&1; echo "set full rc=$?"
ls -la gpurun_out/ev
