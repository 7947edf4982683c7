#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_exact.log 2>&1; echo "exit $?" >> gpurun_out/bench_exact.log
tail -3 gpurun_out/bench_exact.log | cut -c1-3000
timeout 600 python bench.py --steps 10 --warmup 3 --cvf-mode 1 --no-cpu-baseline > gpurun_out/bench_mixed.log 2>&1; echo "exit $?" >> gpurun_out/bench_mixed.log
tail -3 gpurun_out/bench_mixed.log | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "exit $?" >> gpurun_out/bench_reference.log
tail -3 gpurun_out/bench_reference.log | cut -c1-1500
