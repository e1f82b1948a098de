#!/bin/bash
# MC-bound regime A/B: chains per MC warp (1/2/3/4), sleep between poll groups (64 / 250 ns)
for v in "" _c1 _c3 _c4 _s64 _s250; do
  echo "== lib$v"
  M6A_LIB=$PWD/m6anet_b200/libm6anet_b200$v.so timeout 300 python tools/gpu_quick_tc2.py --no-parity --time --only-big --only-tc 2>&1 | grep -E '"encoder"|rror' | cut -c1-100
done
