#!/bin/bash
# round 2, run B: TMEM ring variants: tests, A/B benches, ncu of the default exact and mixed kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py -m gpu -q -x --timeout 600 -s > gpurun_out/pytest_modes.log 2>&1
echo "pytest modes exit $?" >> gpurun_out/pytest_modes.log
tail -5 gpurun_out/pytest_modes.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_modes.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
rm -f gpurun_out/variants_b.txt
for cfg in "0 0" "0 1" "0 2" "0 3" "0 4" "1 0" "1 2" "1 4" "0 0 128" "1 0 128"; do
  set -- $cfg
  extra=""; [ -n "$3" ] && extra="--cta-threads $3"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --cvf-mode $1 --variant $2 $extra 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('mode',j['config']['cvf_mode'],'variant',j['config']['variant'],'$extra','ms/step',round(j['ms_per_step'],3),'cvf_kernel',round(j['roofline']['kernel_ms'],3),'frac',round(j['roofline']['frac'],4),'e2e',round(j['e2e']['ms_per_step'],3))
    else: print(l.rstrip()[:300])
" | tee -a gpurun_out/variants_b.txt
done
for m in 0 1; do
ncu --set full --clock-control none --import-source on -k regex:cvf_stream -s 2 -c 1 -f -o gpurun_out/cvf_prof_b_mode$m \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cvf-mode $m > gpurun_out/bench_under_ncu_mode$m.log 2>&1
done
ls -la gpurun_out | head -30
