#!/bin/bash
for v in "" _c4; do
  echo "== lib$v (ffma)"
  M6A_ENCODER=ffma M6A_LIB=$PWD/m6anet_b200/libm6anet_b200$v.so timeout 300 python tools/gpu_quick_tc2.py --no-parity --time --only-big 2>&1 | grep -E '"encoder": "ffma"|rror' | cut -c1-190
done
