#!/bin/bash
# multi-GPU: multi-process parity test (CUDA IPC exchange) + bench at N = all GPUs of the box (and N=2 when more)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_multiprocess.py -m gpu -q -x --timeout 800 -s > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -6 gpurun_out/pytest_multi.log
run() { # n, extra args
  n=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29400+n)) bench.py --gpus $n --steps 10 --warmup 3 "$@" 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('N',j['n_gpus'],j['config']['workload'][:2],j['config']['cvf_mode'],j['config']['parallelism'],'value',round(j['value'],1),'ms',round(j['ms_per_step'],3),'e2e',round(j['e2e']['value'],1),'e2e_ms',round(j['e2e']['ms_per_step'],3),'parity',j['parity_checked'],j['config']['stage_ms_last_step']); print(l.rstrip())
    else: print(l.rstrip()[:400])
" >> gpurun_out/multi_bench.txt
}
rm -f gpurun_out/multi_bench.txt
for n in 2 4 8; do
  if [ $n -le $N ]; then
    run $n
    run $n --exchange p2p-barrier
    run $n --exchange nccl
    run $n --workload C5
    run $n --cvf-mode 1
  fi
done
grep "^N " gpurun_out/multi_bench.txt
