#!/bin/bash
# runs the kernel-time part of the bench with each prebuilt library variant
cd /root/repo
cp data_compressor_b200/lib/libdc_cuda.so /tmp/lib_orig.so
for f in profiles/variants_tmp/*.so; do
  cp $f data_compressor_b200/lib/libdc_cuda.so
  echo -n "$(basename $f) "
  python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print(round(k['front'],3), round(k['back'],3), d['round_trip_ok'])"
done
cp /tmp/lib_orig.so data_compressor_b200/lib/libdc_cuda.so
