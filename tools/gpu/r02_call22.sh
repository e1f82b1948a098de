#!/bin/bash
# MMA issuer warp-converged with elect.sync (back-to-back UTCHMMA): parity, timing, phase profile
bash tools/gpu/r02_call20.sh
for it in 1 1000; do
  echo "== profile iters=$it"
  M6A_LIB=$PWD/m6anet_b200/libm6anet_b200_prof.so timeout 300 python tools/gpu_tc_profile.py 1000000 $it 2>&1 | tail -32
done
