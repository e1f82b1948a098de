#!/bin/bash
# A/B: D2 handed over per accumulator group (default) vs as a whole; phase profiles of both
for v in "" _nosplit; do
  echo "== lib$v"
  M6A_LIB=$PWD/m6anet_b200/libm6anet_b200$v.so timeout 300 python tools/gpu_quick_tc2.py --no-parity --time --only-big --only-tc 2>&1 | grep -E '"encoder"|rror' | cut -c1-100
done
for v in 1 0; do for it in 1 1000; do
  echo "== profile split=$v iters=$it"
  M6A_LIB=$PWD/m6anet_b200/libm6anet_b200_prof$v.so timeout 300 python tools/gpu_tc_profile.py 1000000 $it 2>&1 | tail -32
done; done
