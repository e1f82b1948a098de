#!/bin/bash
# ncu --set full of the "barrier skeleton + Monte-Carlo only" ablation build: what bounds the pooling when nothing else computes
M6A_LIB=$PWD/m6anet_b200/libm6anet_b200_abl30.so ncu --set full --clock-control none --import-source on -k regex:mil_infer_tc -s 3 -c 1 -o gpurun_out/r02_abl30 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-digest --no-e2e > gpurun_out/r02_abl30_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_abl30_ncu.log
