#!/bin/bash
# D2 handed back per accumulator group + LEA index addressing + balanced tile size: parity, then timing incl. one N=8 shard
timeout 600 python tools/gpu_quick_tc2.py gpurun_out/r02_tc2.json --time --big > gpurun_out/r02_tc2.log 2>&1; echo "tc2 rc=$?"
grep -E "KERNEL|stuck|Error|error" gpurun_out/r02_tc2.log | cut -c1-250 | head -5
grep -E '"read_vs_p32"' gpurun_out/r02_tc2.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[5:]); print(l[:5], d['case'], 'p32 %.2e p64 %.2e site %.2e tc-ffma %.2e mc %d' % (d['read_vs_p32'], d['read_vs_p64'], d['site_vs_oracle'], d['site_tc_vs_ffma'], d['mod_count_diffs']))"
grep '"encoder"' gpurun_out/r02_tc2.log | cut -c1-200
