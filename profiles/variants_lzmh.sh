#!/bin/bash
# LZMH encode / decode kernel times with each prebuilt library variant in profiles/variants_tmp/
cd /root/repo
cp data_compressor_b200/lib/libdc_cuda.so /tmp/lib_orig.so
for f in profiles/variants_tmp/*.so; do
  cp $f data_compressor_b200/lib/libdc_cuda.so
  echo -n "$(basename $f) "
  python bench.py --workload lzmh --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('enc', round(d['ms_per_step'],2), 'dec', round(d['decode_ms'],2), d['blocks_fully_restored'])"
done
cp /tmp/lib_orig.so data_compressor_b200/lib/libdc_cuda.so
