#!/bin/bash
mkdir -p gpurun_out
for v in "" _hint _sleep _sleep5; do
  echo "== lib$v"
  M6A_LIB=$PWD/m6anet_b200/libm6anet_b200$v.so timeout 300 python tools/gpu_quick_tc2.py --no-parity --time --only-big --only-tc 2>&1 | grep '"encoder"' | cut -c1-110
done
