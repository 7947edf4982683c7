#!/bin/bash
# rows-per-segment A/B for the per-rank share of an 8-GPU run (16 slices of C4, 32 slices of C5), emulated on one GPU
mkdir -p gpurun_out
run() {
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" 2>>gpurun_out/bench_x.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$*', 'slices', d['slices'], 'ms %.3f cvf_kernel %.4f'%(d['ms_per_step'],d['cvf_kernel_ms']))
" | tee -a gpurun_out/r2_segrows_shards_ab.txt
}
for sr in 0 216 180 120 90; do run --workload C4 --emulate-shards 8 --seg-rows $sr; done
for sr in 0 216 135 120 90; do run --workload C5 --emulate-shards 8 --seg-rows $sr; done
for sr in 0 270 180; do run --workload C4 --emulate-shards 4 --seg-rows $sr; done
tail -3 gpurun_out/bench_x.err
