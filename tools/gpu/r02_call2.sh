#!/bin/bash
# round-2 GPU call 2: first run of the warp-specialised tensor-core kernel
mkdir -p gpurun_out
timeout 600 python tools/gpu_quick_tc2.py gpurun_out/r02_tc2.json --time --big > gpurun_out/r02_tc2.log 2>&1; echo "tc2 rc=$?"
tail -60 gpurun_out/r02_tc2.log
