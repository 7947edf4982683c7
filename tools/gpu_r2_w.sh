#!/bin/bash
# finer row segments for the streaming kernel (the wave planner's choice is --seg-rows 0): C4 and C3, exact + tolerance mode
mkdir -p gpurun_out
run() {
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity "$@" 2>>gpurun_out/bench_w.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$*', 'value %.1f ms %.3f cvf_kernel %.3f mixed %.1f kernel %.3f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d.get('tolerance_mode',{}).get('value',0),d.get('tolerance_mode',{}).get('kernel_ms',0)))
" | tee -a gpurun_out/r2_segrows_ab2.txt
}
for sr in 0 270 216 180 135 108; do run --workload C4 --seg-rows $sr; done
for sr in 0 90 80 60; do run --workload C3 --seg-rows $sr; done
tail -3 gpurun_out/bench_w.err
