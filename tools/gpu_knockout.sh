#!/bin/bash
# builds tools/cvf_knockout for a list of knockout masks ON THE BOX (nvcc is in the image) and runs them
mkdir -p gpurun_out
rm -f gpurun_out/knockout.txt
for m in 0 1 2 4 8 16 32 5 36 63; do
  nvcc -O3 -std=c++17 -fmad=false -gencode arch=compute_100a,code=sm_100a -DPSM_KNOCKOUT=$m -o /tmp/ko_$m tools/cvf_knockout.cu 2>/dev/null && /tmp/ko_$m | tee -a gpurun_out/knockout.txt
done
echo "# bits: 1 shuffles, 2 TMEM ring, 4 guide loads, 8 stores, 16 stage-1 widening+fp64 running sums, 32 volume loads" >> gpurun_out/knockout.txt
