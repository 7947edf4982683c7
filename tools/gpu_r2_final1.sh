#!/bin/bash
# final single-GPU validation: every GPU test, smoke, the default bench line (exact) + mixed + reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/bench_default.err; echo "exit $?" >> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.err; cut -c1-250 gpurun_out/r2_bench_default.json
python bench.py --cvf-mode 1 --no-cpu-baseline > gpurun_out/r2_bench_mixed.json 2>> gpurun_out/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/bench_default.err
cut -c1-200 gpurun_out/r2_bench_mixed.json; cut -c1-200 gpurun_out/r2_bench_reference.json
