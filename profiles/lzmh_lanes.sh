#!/bin/bash
# LZMH decode time against the number of lanes per warp that take a block (DC_LZMH_DECODE_LANES), 4096 and 16384 blocks
cd "$(dirname "$0")/.."
for blocks in 4096 16384; do
  for lanes in 1 2 4 8 16 32 auto; do
    if [ $lanes = auto ]; then unset DC_LZMH_DECODE_LANES; else export DC_LZMH_DECODE_LANES=$lanes; fi
    echo -n "blocks $blocks lanes $lanes: "
    python bench.py --workload lzmh --blocks $blocks --steps 2 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('decode %.2f ms  %.1f GB/s   (encode %.2f ms)' % (d['decode_ms'], d['decode_gbs'], d['ms_per_step']))"
  done
done
