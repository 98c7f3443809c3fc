#!/bin/bash
# One gpurun call for the decoder-side output conversion:
#   gpurun --timeout 600 -- 'bash scripts/r2_decode_gpu.sh'
# parity tests, timing of every output format, one ncu --set full capture of the RGB32 kernel.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_decode_gpu.py -x -q > gpurun_out/decode_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/decode_pytest.log
tail -15 gpurun_out/decode_pytest.log
timeout 120 python scripts/probe_decode.py 48 10 > gpurun_out/decode_probe.json 2> gpurun_out/decode_probe.err; cat gpurun_out/decode_probe.json; tail -3 gpurun_out/decode_probe.err
timeout 200 ncu --clock-control none --set full --import-source on -k regex:dec_packed_kernel --launch-skip 3 -c 1 -f -o gpurun_out/dec_packed_kernel_r2 \
    python scripts/probe_decode.py 48 2 bgra_bottom_up > gpurun_out/ncu_dec_packed_r2.log 2>&1
tail -3 gpurun_out/ncu_dec_packed_r2.log
ls -la gpurun_out | tail -8
