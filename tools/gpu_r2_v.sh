#!/bin/bash
# CVC slice chunking (DRAM page locality of the store stream) and guide-precompute rows per warp: parity, then stage-time A/B at C4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modes.py -m gpu -q --timeout 500 -k "cvc or guide" > gpurun_out/pytest_v.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_v.log
tail -5 gpurun_out/pytest_v.log
run() {
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity "$@" 2>>gpurun_out/bench_v.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
s=d['config']['stage_ms_last_step']
print('$*', 'value %.1f ms %.3f cvc %.4f guide %.4f'%(d['value'],d['ms_per_step'],s['cvc'],s['cvf']-s['cvf_kernel']))
" | tee -a gpurun_out/r2_cvc_chunk_ab.txt
}
for ch in 0 64 32 16 8 4; do run --cvc-chunk $ch; done
run --cvc-chunk 16 --cvc-variant 2
run --cvc-chunk 16 --cvc-variant 1
for gr in 32 24 16 12 8; do run --guide-rows $gr; done
tail -3 gpurun_out/bench_v.err
