#!/bin/bash
# scaling run the way the driver does it: N = 1,2,4,8 back to back
mkdir -p gpurun_out
for n in ${NLIST:-1 2 4 8}; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_n$n.log 2>&1
  fi
  echo "exit $?" >> gpurun_out/scale_n$n.log
  grep "^{" gpurun_out/scale_n$n.log | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('N',j['n_gpus'],'value',round(j['value'],1),'ms',round(j['ms_per_step'],3),'e2e',round(j['e2e']['value'],1),j['config']['stage_ms_last_step'])" || tail -5 gpurun_out/scale_n$n.log
done
