#!/bin/bash
for v in "" _sl _hint; do
  echo "== lib$v"
  M6A_LIB=$PWD/m6anet_b200/libm6anet_b200$v.so timeout 300 python tools/gpu_quick_tc2.py --no-parity --time --only-big --only-tc 2>&1 | grep -E '"encoder"|rror|KERNEL' | cut -c1-100
done
