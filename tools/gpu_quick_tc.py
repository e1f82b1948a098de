#!/usr/bin/env python
"""Torch-free GPU check + timing of the EXPERIMENTAL tensor-core read encoder (m6anet_b200/csrc/experimental).
    make -C m6anet_b200/csrc/experimental && timeout 120 python tools/gpu_quick_tc.py [out.json] [--time]
Per-read probabilities of libm6a_encoder_tc.so against the oracle on the golden synthetic inputs (a partial tile, one
tile, all 14 108 reads; 3 weight sets), device memory through libcudart + ctypes.  --time: 10 M reads, kernel time by
CUDA events next to the product kernel at 1 iteration (encoder-dominated).  Exits non-zero above 5e-6."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from m6anet_b200 import _cabi                                     # noqa: E402
from conftest import ALL_TAGS, GOLDEN, oracle_params              # noqa: E402
from test_encoder_tc import EXP_LIB, folded                       # noqa: E402
from oracle import read_probabilities                             # noqa: E402  (checker)

L = C.CDLL(EXP_LIB)
_cabi.lib()                                                       # loads libcudart.so.12
rt = C.CDLL("libcudart.so.12")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
rt.cudaFree.argtypes = [C.c_void_p]
rt.cudaGetErrorString.restype = C.c_char_p


def dev(a=None, nbytes=0):
    p = C.c_void_p()
    n = a.nbytes if a is not None else nbytes
    assert rt.cudaMalloc(C.byref(p), max(n, 16)) == 0
    if a is not None and n:
        assert rt.cudaMemcpy(p, a.ctypes.data_as(C.c_void_p), n, 1) == 0
    return p


def run_tc(h, feats, off, kmer):
    R, S = len(feats), len(off) - 1
    d = [dev(np.ascontiguousarray(feats, np.float32)), dev(np.ascontiguousarray(off, np.int64)),
         dev(np.ascontiguousarray(kmer, np.int32)), dev(nbytes=4 * R)]
    rc = L.m6a_tc_read_probs_f32(h, d[0], d[1], d[2], C.c_int64(S), C.c_int64(R), d[3], None)
    assert rc == 0, rc
    e = rt.cudaDeviceSynchronize()
    assert e == 0, rt.cudaGetErrorString(e)
    out = np.empty(R, dtype=np.float32)
    assert rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), d[3], out.nbytes, 2) == 0
    for p in d:
        rt.cudaFree(p)
    return out


z = np.load(os.path.join(GOLDEN, "synthetic_inputs.npz"))
feats, off, kmer = z["feats"], z["read_off"], z["kmer_idx"]
res = {"cases": [], "worst": 0.0}
for tag in ALL_TAGS:
    st, keep, emb, w = folded(tag)
    h = C.c_void_p()
    rc = L.m6a_tc_create(C.byref(st), C.byref(h))
    assert rc == 0, rc
    P = oracle_params(tag)
    for n_sites in (1, 5, len(off) - 1):
        R = int(off[n_sites])
        got = run_tc(h, feats[:R], off[:n_sites + 1], kmer[:n_sites])
        rows = None if P.emb is None else np.repeat(kmer[:n_sites], np.diff(off[:n_sites + 1]), axis=0)
        want = read_probabilities(P, feats[:R], rows)
        dmax = float(np.abs(got - want).max())
        res["cases"].append({"tag": tag, "reads": R, "max_abs_diff": dmax})
        res["worst"] = max(res["worst"], dmax)
        print(f"{tag:24s} reads {R:6d}  max|dp| = {dmax:.3e}", flush=True)
    L.m6a_tc_destroy(h)

# ---- fused kernel (tensor-core encoder + the product's Monte-Carlo phase) against the oracle ----------------------------
from oracle import mil_inference          # noqa: E402  (checker)
L.m6a_tc_mil_infer_f32.argtypes = [C.c_void_p] * 4 + [C.c_int64] * 3 + [C.c_int32] * 2 + [C.c_uint64, C.c_float] + \
                                  [C.c_void_p] * 4 + [C.c_int64, C.c_void_p]


def run_fused(h, feats, off, kmer, n_iters, seed, base, thr):
    R, S = len(feats), len(off) - 1
    ws = (R // 64 + 2) * 8
    d = [dev(np.ascontiguousarray(feats, np.float32)), dev(np.ascontiguousarray(off, np.int64)),
         dev(np.ascontiguousarray(kmer, np.int32)), dev(nbytes=4 * R), dev(nbytes=4 * S), dev(nbytes=4 * S), dev(nbytes=ws)]
    rc = L.m6a_tc_mil_infer_f32(h, d[0], d[1], d[2], S, R, base, 20, n_iters, seed, thr, d[3], d[4], d[5], d[6], ws, None)
    assert rc == 0, rc
    e = rt.cudaDeviceSynchronize()
    assert e == 0, rt.cudaGetErrorString(e)
    outs = [np.empty(R, np.float32), np.empty(S, np.float32), np.empty(S, np.int32)]
    for o, p in zip(outs, d[3:6]):
        assert rt.cudaMemcpy(o.ctypes.data_as(C.c_void_p), p, o.nbytes, 2) == 0
    for p in d:
        rt.cudaFree(p)
    return outs


if "--no-fused" not in sys.argv:
    tag = "HCT116_RNA002"
    st, keep, emb, w = folded(tag)
    h = C.c_void_p()
    assert L.m6a_tc_create(C.byref(st), C.byref(h)) == 0
    base = 7_000_000_000
    rp, sp, mc = run_fused(h, feats, off, kmer, 200, 1234, base, 0.033379376)
    orp, osp, omc = mil_inference(oracle_params(tag), feats, off, kmer, n_iters=200, seed=1234, site_id_base=base)
    d_read, d_site = float(np.abs(rp - orp).max()), float(np.abs(sp - osp).max())
    print(f"fused kernel: max|dp_read| = {d_read:.3e}  max|dp_site| = {d_site:.3e}  mod_count diffs = {int((mc != omc).sum())}", flush=True)
    res["fused"] = {"max_abs_diff_read": d_read, "max_abs_diff_site": d_site, "mod_count_diffs": int((mc != omc).sum())}
    res["worst"] = max(res["worst"], d_read)
    if d_site > 1e-4:
        res["worst"] = 1.0
    L.m6a_tc_destroy(h)

if "--time" in sys.argv:
    from m6anet_b200 import weights as W
    from m6anet_b200.engine import MilEngine
    rng = np.random.default_rng(0)
    S, n = 200_000, 50
    tf = rng.standard_normal((S * n, 9), dtype=np.float32)
    to = np.arange(S + 1, dtype=np.int64) * n
    tk = kmer[rng.integers(0, len(kmer), S)].astype(np.int32)
    st, keep, emb, w = folded("HCT116_RNA002")
    h = C.c_void_p()
    assert L.m6a_tc_create(C.byref(st), C.byref(h)) == 0
    d = [dev(tf), dev(to), dev(tk), dev(nbytes=4 * S * n)]
    ev = [C.c_void_p(), C.c_void_p()]
    for e in ev:
        assert rt.cudaEventCreate(C.byref(e)) == 0
    ms = C.c_float()
    best = 1e9
    for it in range(6):
        rt.cudaEventRecord(ev[0], None)
        assert L.m6a_tc_read_probs_f32(h, d[0], d[1], d[2], C.c_int64(S), C.c_int64(S * n), d[3], None) == 0
        rt.cudaEventRecord(ev[1], None)
        assert rt.cudaEventSynchronize(ev[1]) == 0
        rt.cudaEventElapsedTime(C.byref(ms), ev[0], ev[1])
        if it:
            best = min(best, ms.value)
    res["tc_encoder_ms_10M_reads"] = best
    print(f"tensor-core encoder: {best:.3f} ms for {S * n} reads ({S * n / best / 1e6:.1f} G reads/s)")
    eng = MilEngine(W.from_npz(os.path.join(ROOT, "m6anet_b200", "assets", "model_states", "rna002_hct116.npz")), 0)
    eng.infer_host(tf, to, tk, 1, seed=0)
    t = time.perf_counter()
    eng.infer_host(tf, to, tk, 1, seed=0)
    res["product_infer_host_1_iteration_ms"] = (time.perf_counter() - t) * 1e3
    print(f"product kernel through host buffers, 1 iteration: {res['product_infer_host_1_iteration_ms']:.1f} ms (includes PCIe)")
    # fused vs product, device-resident, 1000 iterations, CUDA events
    PL = _cabi.lib()
    R = S * n
    ws = max(int(PL.m6a_mil_workspace_bytes(R)), (R // 64 + 2) * 8)
    o = [dev(nbytes=4 * R), dev(nbytes=4 * S), dev(nbytes=4 * S), dev(nbytes=ws)]

    def timed(launch):
        best = 1e9
        for it in range(5):
            rt.cudaEventRecord(ev[0], None)
            assert launch() == 0
            rt.cudaEventRecord(ev[1], None)
            assert rt.cudaEventSynchronize(ev[1]) == 0
            rt.cudaEventElapsedTime(C.byref(ms), ev[0], ev[1])
            if it:
                best = min(best, ms.value)
        return best

    t_fused = timed(lambda: L.m6a_tc_mil_infer_f32(h, d[0], d[1], d[2], S, R, 0, 20, 1000, 0, 0.033379376, o[0], o[1], o[2], o[3],
                                                  ws, None))
    t_prod = timed(lambda: PL.m6a_mil_infer_f32(eng._handle, d[0], d[1], d[2], S, R, 0, 20, 1000, 0, None, 0.033379376, o[0],
                                                o[1], o[2], o[3], ws, None))
    res["fused_ms_200k_sites"] = t_fused
    res["product_ms_200k_sites"] = t_prod
    print(f"200k sites x 50 reads x 1000 iterations, device-resident: fused tensor-core kernel {t_fused:.3f} ms, "
          f"product kernel {t_prod:.3f} ms")

res["ok"] = res["worst"] <= 5e-6
out_path = next((x for x in sys.argv[1:] if not x.startswith("--")), None)
if out_path:
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump(res, fh, indent=1)
print("OK" if res["ok"] else "FAIL", f"worst {res['worst']:.3e}")
sys.exit(0 if res["ok"] else 1)
