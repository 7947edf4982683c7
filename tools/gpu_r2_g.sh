#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_modes.log 2>&1
echo "pytest modes exit $?" >> gpurun_out/pytest_modes.log
tail -5 gpurun_out/pytest_modes.log
rm -f gpurun_out/variants_g.txt
for cfg in "0 0" "0 9" "1 0" "1 9" "0 9 128" "1 9 128"; do
  set -- $cfg
  extra=""; [ -n "$3" ] && extra="--cta-threads $3"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --cvf-mode $1 --variant $2 $extra 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('mode',j['config']['cvf_mode'],'variant',j['config']['variant'],'$extra','ms/step',round(j['ms_per_step'],3),'cvf_kernel',round(j['roofline']['kernel_ms'],3),'frac',round(j['roofline']['frac'],4),'e2e',round(j['e2e']['ms_per_step'],3),'parity',j['parity_checked'], j['config']['stage_ms_last_step'])
    else: print(l.rstrip()[:300])
" | tee -a gpurun_out/variants_g.txt
done
ncu --set full --clock-control none --import-source on -k regex:cvf_stream -s 2 -c 1 -f -o gpurun_out/cvf_prof_g_mode1 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --cvf-mode 1 --variant 9 > gpurun_out/bench_under_ncu_mode1.log 2>&1
