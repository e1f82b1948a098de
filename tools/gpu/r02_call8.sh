#!/bin/bash
# ncu --set full of the tensor-core kernel (200k sites x 50 x 1000): 3 launches after 2 warm-ups
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mil_infer_tc -s 1 -c 1 -o gpurun_out/r02_tc_v3 -f \
  python tools/gpu_quick_tc2.py --no-parity --time --only-tc > gpurun_out/r02_ncu_v3.log 2>&1
echo rc=$?; tail -3 gpurun_out/r02_ncu_v3.log; ls -la gpurun_out/*.ncu-rep
