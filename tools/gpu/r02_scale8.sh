#!/bin/bash
# 8-GPU box: concurrent H2D ceiling, then the headline bench at N = 1, 2, 4, 8 (digests must agree), numactl facts
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > gpurun_out/r02_lscpu.txt
# (the concurrent H2D ceiling of this box type was measured by an earlier run of this script: profiles/r02_h2d_ceiling.json)
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err; echo "n1 rc=$?"
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err; echo "n$n rc=$?"
done
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/r02_scale_n{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value']/1e6,2), 'M sites/s', round(d['ms_per_step'],3), 'ms; kernel', d['roofline']['kernel_ms'], 'e2e', round(d['e2e']['value']/1e6,2), d['e2e'].get('h2d_gbs'), d['e2e'].get('h2d_ceiling_gbs_per_gpu'), d['parity']['ok'], d['result_digest']['site_prob_mod_count_sha256'][:16], (d['result_digest']['read_prob_sha256'] or '')[:16], d['roofline'].get('per_rank'))
    except Exception as e: print(n, 'ERR', e)
PY
