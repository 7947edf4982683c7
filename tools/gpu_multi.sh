#!/bin/bash
N=${N:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
  fi
  echo "exit $?" >> gpurun_out/bench_n$n.log
  tail -2 gpurun_out/bench_n$n.log | cut -c1-1200
done
