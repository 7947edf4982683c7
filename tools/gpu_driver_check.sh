#!/bin/bash
# what the driver runs at round end: smoke(), bench (ours + reference arm), gpu tests
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
( time python bench.py ) > gpurun_out/bench_default.log 2>&1; echo "exit $?" >> gpurun_out/bench_default.log; tail -5 gpurun_out/bench_default.log | cut -c1-2500
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/bench_reference.log 2>&1; echo "exit $?" >> gpurun_out/bench_reference.log; tail -5 gpurun_out/bench_reference.log | cut -c1-1200
