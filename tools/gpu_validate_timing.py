#!/usr/bin/env python
"""Timing of the validate()-style literal MIL forward: m6a_mil_validate_host_f32 on a labelled
synthetic set against the reference's own validate() loop (utils/training_utils.py:236-256, oracle/_ref) on a bounded
sample of the same sites.  Prints one JSON line.     python tools/gpu_validate_timing.py [sites] [reads] [passes]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m6anet_b200 import weights as W                      # noqa: E402
from m6anet_b200.engine import MilEngine                  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rng = np.random.default_rng(0)
feats = rng.standard_normal((S * n, 9), dtype=np.float32)
off = np.arange(S + 1, dtype=np.int64) * n
kmer = rng.integers(0, 66, size=(S, 3)).astype(np.int32)
eng = MilEngine(W.from_npz(os.path.join(ROOT, "m6anet_b200", "assets", "model_states", "rna002_hct116.npz")), 0)
eng.validate_host(feats[: 50 * n], off[:51], kmer[:50], passes, seed=0)            # warm-up
t0 = time.perf_counter()
rp, bag, mean, mc = eng.validate_host(feats, off, kmer, passes, seed=0, pooling="prod")
dt = time.perf_counter() - t0
line = {"what": "validate()-style literal MIL forward: bags of 20 reads without replacement, one per (site, pass)",
        "sites": S, "reads_per_site": n, "passes": passes, "gpu_seconds_host_buffers": dt, "gpu_sites_per_s": S / dt,
        "gpu_bags_per_s": S * passes / dt, "kernel": "mil_infer_kernel<20, true> (FFMA2 encoder + bag pooling)",
        "launch": eng.last_launch()}
try:
    from oracle import ref_arm                            # checker / baseline only
    if ref_arm.available():
        cores = len(os.sched_getaffinity(0))
        n_ref = min(S, 4000)
        r = ref_arm.time_stock_validate("HCT116_RNA002", feats[: n_ref * n], off[: n_ref + 1], kmer[:n_ref], passes, cores)
        line["reference_validate"] = {"sites_per_s": r["sites_per_s"], "sample_sites": n_ref, "cores": cores,
                                      "seconds": r["seconds"], "how": "the reference's validate() on resident inputs (oracle/_ref)"}
        line["speedup_vs_reference_loop"] = (S / dt) / r["sites_per_s"]
except Exception as exc:                                  # noqa: BLE001
    line["reference_validate"] = {"error": repr(exc)}
print(json.dumps(line))
