#!/bin/bash
# usage: variants_run.sh NAME...  -- swaps data_compressor_b200/lib/variants/NAME/libdc_cuda.so in, checks parity on the coder tests, prints kernel times
mkdir -p gpurun_out/var
cp data_compressor_b200/lib/libdc_cuda.so /tmp/orig.so
for v in base "$@"; do
  if [ "$v" != base ]; then cp data_compressor_b200/lib/variants/$v/libdc_cuda.so data_compressor_b200/lib/libdc_cuda.so; fi
  t=$(python -m pytest tests/test_gpu_parity.py -x -q -k "bac or golden or garbage or random_stage or batch or full_size or long_code or make_test" 2>&1 | tail -1)
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/var/$v.json 2> gpurun_out/var/$v.err
  python -c "
import json; d=json.load(open('gpurun_out/var/$v.json')); k=d['kernel_ms']; print('$v', '| $t |', round(d['value'],1), d['round_trip_ok'], 'enc %.3f dec %.3f front %.3f back %.3f' % (k['bac_encode'], k['bac_decode'], k['front'], k['back']))"
done
cp /tmp/orig.so data_compressor_b200/lib/libdc_cuda.so
