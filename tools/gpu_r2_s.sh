#!/bin/bash
# CVC with grouped 128-bit window loads: parity of the three builds, then stage time A/B on the C4 frame (and at 16 slices = one of 8 ranks)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py -m gpu -q --timeout 600 > gpurun_out/pytest_s.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_s.log
tail -8 gpurun_out/pytest_s.log
for v in 0 1 2; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --cvc-variant $v 2>>gpurun_out/bench_s.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('cvc_variant $v value %.1f ms %.3f'%(d['value'],d['ms_per_step']), d['config']['stage_ms_last_step'])
" | tee -a gpurun_out/r2_cvc_ab.txt
done
tail -3 gpurun_out/bench_s.err
