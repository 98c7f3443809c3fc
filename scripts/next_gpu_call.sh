#!/bin/bash
# Everything that changed after the round's GPU minutes ran out, in ONE gpurun call (about 6 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/next_gpu_call.sh'
# 1. the GPU parity suite (the hpel kernel was reworked on the CPU lockstep simulation only: 252 -> 182
#    instructions per row), 2. its timing at the BASELINE sizes, 3. one ncu --set full capture of
#    it alone, 4. the default bench line.  Results land in gpurun_out/ (summaries: scripts/summarize_profiles.py).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/next_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/next_pytest_gpu.log
timeout 120 python scripts/probe_hpel.py > gpurun_out/next_probe_hpel.json 2> gpurun_out/next_probe_hpel.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:hpel_kernel -c 1 -f \
    -o gpurun_out/next_hpel python scripts/probe_hpel.py --once > gpurun_out/next_ncu_hpel.log 2>&1
timeout 400 python bench.py > gpurun_out/next_bench_n1.json 2> gpurun_out/next_bench_n1.err
tail -3 gpurun_out/next_pytest_gpu.log; cat gpurun_out/next_probe_hpel.json
