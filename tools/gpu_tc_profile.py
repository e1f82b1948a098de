#!/usr/bin/env python
"""Phase profile of mil_infer_tc_kernel (library built with -DM6A_TC_PROFILE=1, selected with M6A_LIB):
cycles per phase of one encoder thread, the MMA issuer and one Monte-Carlo thread of block 0.
    M6A_LIB=m6anet_b200/libm6anet_b200_prof.so python tools/gpu_tc_profile.py [sites] [iters]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m6anet_b200 import _cabi, weights as W      # noqa: E402
from m6anet_b200.engine import MilEngine         # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
n = 50
eng = MilEngine(W.from_npz(os.path.join(ROOT, "m6anet_b200", "assets", "model_states", "rna002_hct116.npz")), 0)
eng.set_encoder("tc")
rng = np.random.default_rng(0)
f = rng.standard_normal((S * n, 9), dtype=np.float32)
o = np.arange(S + 1, dtype=np.int64) * n
k = rng.integers(0, 66, size=(S, 3)).astype(np.int32)
eng.infer_host(f, o, k, iters, n_chunks=1)
eng.infer_host(f, o, k, iters, n_chunks=1)
L = _cabi.lib()
out = (C.c_uint64 * 40)()
L.m6a_debug_tc_profile.argtypes = [C.c_void_p]
assert L.m6a_debug_tc_profile(out) == 0
v = list(out)
tiles = S * n / 128 / 148
roles = [
    ("staging + E2 group 0 row 0 (every second tile)", 0, ["wait l1_done(t-2)", "cp.async wait", "split + STS", "fence.proxy.async", "arrive x_full", "next_tile", "prefetch issue",
                                   "wait d2_full", "E2 TMEM read + logits", "E2 sigmoid + outputs"]),
    ("MMA issuer", 10, ["wait x_full", "L1 issue+commit", "wait d2_free", "wait a_full (5x)", "L2 issue+commit (5x)"]),
    ("E1 thread 0", 30, ["wait l1_done", "tmem_ld+wait (5x)", "relu/split (5x)", "wait a_free (5x)", "tmem_st+wait+arrive (5x)"]),
    ("MC warp 0", 20, ["wait slab_full", "sites of the slab", "-", "report"]),
]
print(f"sites {S} iters {iters}: ~{tiles:.0f} MMA tiles per CTA; cycles per tile (block 0):")
for name, base, phases in roles:
    print(f" {name}: total {sum(v[base:base + 10]) / tiles:.0f}")
    for i, nm in enumerate(phases):
        print(f"   {nm:28s} {v[base + i] / tiles:9.0f}")
