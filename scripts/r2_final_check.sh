#!/bin/bash
# What the driver runs at round end, in one gpurun call: gpurun --timeout 1500 -- 'bash scripts/r2_final_check.sh'
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu.log
tail -4 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/final_bench.json; tail -3 gpurun_out/final_bench.err
