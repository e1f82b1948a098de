#!/bin/bash
mkdir -p gpurun_out
export M6A_LIB=$PWD/m6anet_b200/libm6anet_b200_abl30.so
python tools/gpu_quick_tc2.py --no-parity --time --only-big --only-tc 2>&1 | grep '"encoder"' | cut -c1-100
ncu --set full --clock-control none --import-source on -k regex:mil_infer_tc -s 1 -c 1 -o gpurun_out/r02_tc_abl30 -f \
  python tools/gpu_quick_tc2.py --no-parity --time --only-tc > gpurun_out/r02_ncu_abl30.log 2>&1
echo rc=$?
