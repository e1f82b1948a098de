#!/usr/bin/env python
"""Torch-free GPU check of the validate()-style entry points against the oracle (seconds, not minutes: no torch import).
    python tools/gpu_quick_validate.py [out.json] [--time]
Runs m6a_mil_validate_host_f32 (Floyd bags / inference stream, prod|mean|max, bag sizes 20 and 7) on the 288 synthetic
golden sites and m6a_mil_validate_f32 with explicit bags (the reference's replayed MT19937 draws) on the bundled sets,
device memory through libcudart + ctypes.  Prints max abs differences; exits non-zero above 2e-6."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m6anet_b200 import _cabi, weights as W          # noqa: E402
from m6anet_b200.engine import MilEngine, POOLING    # noqa: E402
from oracle import ReadEncoderParams, mil_validate, sample_indices_many   # noqa: E402  (checker)

t0 = time.time()
G = os.path.join(ROOT, "tests", "golden")
NPZ = os.path.join(ROOT, "m6anet_b200", "assets", "model_states", "rna002_hct116.npz")
eng = MilEngine(W.from_npz(NPZ), 0)
P = ReadEncoderParams.from_npz(NPZ)
si = np.load(os.path.join(G, "synthetic_inputs.npz"))
feats, off, kmer = si["feats"], si["read_off"], si["kmer_idx"]
res = {"cases": [], "worst": 0.0}


def note(name, *pairs):
    d = max(float(np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64)))) for a, b in pairs)
    nan_ok = all(np.array_equal(np.isnan(a), np.isnan(b)) for a, b in pairs)
    res["cases"].append({"case": name, "max_abs_diff": d, "nan_pattern_equal": nan_ok})
    res["worst"] = max(res["worst"], d if nan_ok else 1.0)
    print(f"{name:48s} max|d| = {d:.3e} nan_ok={nan_ok}", flush=True)


base = 7_000_000_000
for pool in ("prod", "mean", "max"):
    for k, it, rep in ((20, 5, False), (20, 300, False), (7, 40, False), (20, 64, True)):
        rp, bag, mean, mc = eng.validate_host(feats, off, kmer, it, seed=11, site_id_base=base, n_samples=k, pooling=pool,
                                              replace=rep, n_chunks=3)
        if rep:
            idx = sample_indices_many(11, base + np.arange(len(off) - 1), np.diff(off), it, k)
            orp, obag, omean, omc = mil_validate(P, feats, off, kmer, it, n_samples=k, pool=pool, sample_idx=idx)
        else:
            orp, obag, omean, omc = mil_validate(P, feats, off, kmer, it, seed=11, site_id_base=base, n_samples=k, pool=pool)
        # the device averages the passes pairwise (butterfly over lanes, then blocks); hold it to the float64 mean of the
        # oracle's bags -- the reference-order float32 average is formed on the host from bag_prob (validation.py)
        note(f"host {pool} k={k} iters={it} replace={rep}", (rp, orp), (bag, obag), (mean, obag.astype(np.float64).mean(axis=1)))

# explicit bags through the device entry point (libcudart via ctypes)
rt = C.CDLL("libcudart.so.12")   # the runtime libm6anet_b200.so already loaded (same soname)
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
rt.cudaFree.argtypes = [C.c_void_p]


def dev(a=None, nbytes=0):
    p = C.c_void_p()
    n = a.nbytes if a is not None else nbytes
    assert rt.cudaMalloc(C.byref(p), max(n, 16)) == 0
    if a is not None and n:
        assert rt.cudaMemcpy(p, a.ctypes.data_as(C.c_void_p), n, 1) == 0
    return p


def back(p, shape, dtype):
    out = np.empty(shape, dtype=dtype)
    assert rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), p, out.nbytes, 2) == 0
    return out


vg = np.load(os.path.join(G, "validate_golden.npz"))
bf = np.load(os.path.join(G, "bundled_flat.npz"))
L = _cabi.lib()
for pool in ("prod", "mean", "max"):
    for mode in ("Val", "Test"):
        sidx = vg[f"{mode}_site_index"]
        ro = bf["read_off"]
        n = np.diff(ro)[sidx]
        o = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
        f = np.ascontiguousarray(np.concatenate([bf["feats"][ro[s]:ro[s + 1]] for s in sidx]), dtype=np.float32)
        km = np.ascontiguousarray(bf["kmer_idx"][sidx], dtype=np.int32)
        bags = np.ascontiguousarray(vg[f"{mode}_bags"], dtype=np.uint16)
        S, R, it = len(sidx), int(o[-1]), int(vg["n_iters"])
        ws = int(L.m6a_mil_workspace_bytes(R))
        d = [dev(f), dev(o), dev(km), dev(bags), dev(nbytes=4 * R), dev(nbytes=4 * S * it), dev(nbytes=4 * S), dev(nbytes=4 * S),
             dev(nbytes=ws)]
        rc = L.m6a_mil_validate_f32(eng._handle, d[0], d[1], d[2], S, R, 0, 20, it, 0, d[3], POOLING[pool], 0, 0.033379376,
                                    d[4], d[5], d[6], d[7], d[8], ws, None)
        assert rc == 0, rc
        assert rt.cudaDeviceSynchronize() == 0
        bag = back(d[5], (S, it), np.float32)
        note(f"device explicit bags {pool} {mode} vs reference validate()", (bag.T, vg[f"{pool}_{mode}_y_pred"]))
        for p in d:
            rt.cudaFree(p)
# ---- timing (optional: `--time`): validate()-style passes next to the inference pass on the same synthetic sites -------
if "--time" in sys.argv:
    rng = np.random.default_rng(0)
    S, n = 200_000, 50
    tf = rng.standard_normal((S * n, 9), dtype=np.float32)
    to = (np.arange(S + 1, dtype=np.int64) * n)
    tk = kmer[rng.integers(0, len(kmer), S)]
    timing = {}
    for name, fn in (("validate_host 5 passes, bags of 20 without replacement", lambda: eng.validate_host(tf, to, tk, 5, seed=1)),
                     ("validate_host 1000 passes", lambda: eng.validate_host(tf, to, tk, 1000, seed=1)),
                     ("infer_host 1000 iterations", lambda: eng.infer_host(tf, to, tk, 1000, seed=1))):
        fn()
        best = min((lambda t: (fn(), time.perf_counter() - t)[1])(time.perf_counter()) for _ in range(3))
        timing[name] = {"ms": best * 1e3, "sites_per_s": S / best}
        print(f"{name:60s} {best * 1e3:8.1f} ms  {S / best / 1e6:7.2f} M sites/s (host buffers, {S} sites x {n} reads)", flush=True)
    res["timing"] = timing
res["seconds"] = time.time() - t0
res["ok"] = res["worst"] <= 2e-6
out_path = next((x for x in sys.argv[1:] if not x.startswith("--")), None)
if out_path:
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump(res, fh, indent=1)
print("OK" if res["ok"] else "FAIL", f"worst {res['worst']:.3e} in {res['seconds']:.1f} s")
sys.exit(0 if res["ok"] else 1)
