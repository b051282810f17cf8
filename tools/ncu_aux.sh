#!/bin/bash
# ncu full captures of the HBM-bound helper kernels of the GEMM path (one launch each)
TAG=${1:-r01}
for k in split_a_kernel split_b_kernel crt_fast_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$k \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_$k.log 2>&1
done
ls -la gpurun_out/*prof_*
