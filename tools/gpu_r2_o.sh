#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pp.py tests/test_gpu_modes.py -m gpu -q -x --timeout 600 -k "pp or post or async or selection" > gpurun_out/pytest_o.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_o.log
tail -4 gpurun_out/pytest_o.log
python tools/pp_time.py C4 2>&1 | tail -2 | tee gpurun_out/pp_time.txt
python tools/pp_time.py C3 2>&1 | tail -2 | tee -a gpurun_out/pp_time.txt
