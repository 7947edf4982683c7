#!/bin/bash
# iteration loop on the GPU: parity tests (fail fast), bench, optional ncu of the CVF kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-1500
if [ -n "$NCU" ]; then
  ncu --set full --clock-control none --import-source on -k regex:cvf_stream -s 2 -c 1 -f -o gpurun_out/cvf_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
fi
