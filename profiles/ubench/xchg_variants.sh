# Size-exchange variants on two GPUs, side by side on one box.  DC_XCHG / DC_ORDER / DC_RESOLVE were temporary switches
# patched into bench.py and shard.py for this run only (simple = the tree's exchange_sizes_async, ring = preallocated pinned slots on a side stream;
# first / after = collective started before / after the decode is queued; each / all = results fetched one by one / stacked on the device).
# The tree keeps what won: simple, first, each.
mkdir -p gpurun_out/mg
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra"
for rep in 1 2; do
for cfg in "simple first each" "simple first all" "ring first all" "ring after all" "simple after all"; do
  set -- $cfg
  DC_XCHG=$1 DC_ORDER=$2 DC_RESOLVE=$3 $TR > gpurun_out/mg/x.json 2> gpurun_out/mg/x.err
  python -c "
import json; d=json.loads(open('gpurun_out/mg/x.json').read().strip().splitlines()[-1]); print('$cfg', 'ms %.3f' % d['ms_per_step'], 'value %.1f' % d['value'])"
done; done
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('single', 'ms %.3f' % d['ms_per_step'], 'value %.1f' % d['value'])"
