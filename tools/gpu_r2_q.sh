#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_q.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_q.log
tail -25 gpurun_out/pytest_q.log
