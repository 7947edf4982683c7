#!/bin/bash
# the planner refactor (psm_cvf_plan) changes no kernel; whole GPU suite + default bench as the last check of the shipped library
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 250 > gpurun_out/pytest_z.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_z.log
tail -3 gpurun_out/pytest_z.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_z.err | cut -c1-120
