#!/bin/bash
# last sanity pass with the final library: whole GPU suite, the default bench line (CPU arm skipped: measured in gpu_r2_final3.sh), per-rank shares
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 500 > gpurun_out/pytest_y.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_y.log
tail -3 gpurun_out/pytest_y.log
python bench.py --no-cpu-baseline > gpurun_out/r2_bench_final_nocpu.json 2>gpurun_out/bench_y.err; echo "exit $?" >> gpurun_out/bench_y.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_final_nocpu.json'))
print('value %.1f ms %.3f e2e %.1f kernel %.3f frac %.4f mixed %.1f (%.3f) parity %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['tolerance_mode']['value'],d['tolerance_mode']['kernel_ms'],d['parity_checked']), d['config']['stage_ms_last_step'])
"
python bench.py --steps 20 --warmup 5 --workload C4 --emulate-shards 8 2>>gpurun_out/bench_y.err | cut -c150-420
python bench.py --steps 20 --warmup 5 --workload C5 --emulate-shards 8 2>>gpurun_out/bench_y.err | cut -c150-420
tail -2 gpurun_out/bench_y.err
