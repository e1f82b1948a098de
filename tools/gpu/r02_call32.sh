#!/bin/bash
# compute-sanitizer memcheck (and, in a second run, synccheck) of the shipped kernels on the edge-case tests (tiny, ragged, empty, huge sites; both encoders)
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 7 --launch-timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "edge_cases or short_jobs" 2>&1 | tail -6
echo "memcheck rc=${PIPESTATUS[0]}"
