#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests -m gpu -q --timeout 800 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -4 gpurun_out/pytest_all.log
rm -f gpurun_out/multi_bench_n4.txt
run() { n=$1; shift
  if [ $n -eq 1 ]; then cmd="python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline $*";
  else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29400+n)) bench.py --gpus $n --steps 10 --warmup 3 $*"; fi
  timeout 600 $cmd 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('N',j['n_gpus'],j['config']['workload'][:2],j['config']['cvf_mode'],j['config']['parallelism'],'value',round(j['value'],1),'ms',round(j['ms_per_step'],3),'e2e',round(j['e2e']['value'],1),'e2e_ms',round(j['e2e']['ms_per_step'],3),'parity',j['parity_checked'],j['config']['stage_ms_last_step'])
    else: print(l.rstrip()[:400])
" >> gpurun_out/multi_bench_n4.txt
}
run 1; run 2; run $N; run $N --exchange p2p-barrier
grep "^N \|rror" gpurun_out/multi_bench_n4.txt | cut -c1-420
