#!/bin/bash
# final multi-GPU evidence: 8-process parity test (CUDA IPC exchange, device-side flags), bench at N = 8 (C4, C5, barrier variant), then
# N = 4, 2, 1 on the same box for a consistent scaling table
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 600 python -m pytest tests/test_gpu_multiprocess.py -m gpu -q -x --timeout 500 -k "8" > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -4 gpurun_out/pytest_multi.log
rm -f gpurun_out/multi_bench.txt gpurun_out/multi_lines.jsonl
run() { # n, extra args
  n=$1; shift
  if [ $n -eq 1 ]; then cmd="python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline $*";
  else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29400+n)) bench.py --gpus $n --steps 10 --warmup 3 $*"; fi
  timeout 400 $cmd 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('N',j['n_gpus'],j['config']['workload'][:2],j['config']['cvf_mode'],j['config']['parallelism'],'value',round(j['value'],1),'ms',round(j['ms_per_step'],3),'e2e',round(j['e2e']['value'],1),'e2e_ms',round(j['e2e']['ms_per_step'],3),'parity',j['parity_checked'],j['config']['stage_ms_last_step']); open('gpurun_out/multi_lines.jsonl','a').write(l)
    else: print(l.rstrip()[:400])
" >> gpurun_out/multi_bench.txt
}
run $N
run $N --workload C5
run $N --exchange p2p-barrier
run $N --cvf-mode 1
if [ $N -ge 8 ]; then run 4; run 2; fi
run 1
grep "^N \|rror" gpurun_out/multi_bench.txt | cut -c1-420
