#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (+ reference arm), ncu launch list and one full capture of the top kernel.
# usage: tools/gpu_round.sh [tag]   -> artefacts under gpurun_out/<tag>_*
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --extras 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tee gpurun_out/${TAG}_bench_ref.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "== ncu full (top kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 2 -f -o gpurun_out/${TAG}_prof_gemm \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -15
