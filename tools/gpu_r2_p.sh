#!/bin/bash
# packed remainder strips + integer-widening guide kernel: full GPU suite, then A/B timing and the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_p.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_p.log
tail -15 gpurun_out/pytest_p.log
for np in 1 0; do for m in 0 1; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --no-pack $np --cvf-mode $m 2>>gpurun_out/bench_p.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('no_pack',$np,'mode',$m,'value %.1f ms %.3f cvf_kernel %.3f e2e %.1f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['e2e']['value']), d['config']['stage_ms_last_step'])
" | tee -a gpurun_out/r2_pack_ab.txt
done; done
python bench.py --workload C3 --steps 20 --warmup 5 --no-cpu-baseline --no-pack 1 2>>gpurun_out/bench_p.err | cut -c1-120 | tee -a gpurun_out/r2_pack_ab.txt
python bench.py --workload C3 --steps 20 --warmup 5 --no-cpu-baseline 2>>gpurun_out/bench_p.err | cut -c1-120 | tee -a gpurun_out/r2_pack_ab.txt
python bench.py > gpurun_out/r2_bench_default_p.json 2>>gpurun_out/bench_p.err; echo "exit $?" >> gpurun_out/bench_p.err
cut -c1-400 gpurun_out/r2_bench_default_p.json; tail -3 gpurun_out/bench_p.err
