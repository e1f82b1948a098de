#!/bin/bash
# lanes = sites pooling (conflict-free q loads) in the MC-bound regime
for v in "" _ls; do
  echo "== lib$v"
  M6A_LIB=$PWD/m6anet_b200/libm6anet_b200$v.so timeout 300 python tools/gpu_quick_tc2.py --first-only --time --only-big --only-tc 2>&1 | grep -E '"encoder"|rror|^ok|^FAIL' | cut -c1-160
done
