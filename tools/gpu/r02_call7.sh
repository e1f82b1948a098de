#!/bin/bash
mkdir -p gpurun_out
for v in "" _abl1 _abl2 _abl4 _abl8 _abl16 _abl6 _abl30 _abl32; do
  echo "== lib$v"
  M6A_LIB=$PWD/m6anet_b200/libm6anet_b200$v.so timeout 300 python tools/gpu_quick_tc2.py --no-parity --time --only-big --only-tc 2>&1 | grep -E '"encoder"|rror' | cut -c1-100
done
