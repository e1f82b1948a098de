#!/bin/bash
# ablations (results wrong by construction, timing only): 30 = barrier skeleton + MC only; 6 = no E1 math, no MMAs; 2 = no E1 math; 8 = no E2 read-out
for v in _abl30 _abl6 _abl2 _abl8; do
  echo "== lib$v"
  M6A_LIB=$PWD/m6anet_b200/libm6anet_b200$v.so timeout 300 python tools/gpu_quick_tc2.py --no-parity --time --only-big --only-tc 2>&1 | grep -E '"encoder"|rror' | cut -c1-100
done
