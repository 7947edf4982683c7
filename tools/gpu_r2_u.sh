#!/bin/bash
# CVC interior fast path: parity (all builds, every GPU parity test), then stage-time A/B on the C4 frame
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_u.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_u.log
tail -8 gpurun_out/pytest_u.log
for v in 2 0 1; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --cvc-variant $v 2>>gpurun_out/bench_u.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('cvc_variant $v value %.1f ms %.3f mixed %.1f'%(d['value'],d['ms_per_step'],d['tolerance_mode']['value']), d['config']['stage_ms_last_step'])
" | tee -a gpurun_out/r2_cvc_fast_ab.txt
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload C3 2>>gpurun_out/bench_u.err | cut -c1-100 | tee -a gpurun_out/r2_cvc_fast_ab.txt
tail -3 gpurun_out/bench_u.err
