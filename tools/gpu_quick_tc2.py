#!/usr/bin/env python
"""Torch-free GPU check + timing of the tensor-core kernel (mil_infer_tc_kernel) against the oracle and the FFMA kernel.
    python tools/gpu_quick_tc2.py [out.json] [--time] [--big]
Parity: golden synthetic inputs (288 ragged sites, 14 108 reads) for every weight set -- per-read p against the float32
oracle and the float64 truth, site probabilities on the shared index stream, mod_count; then edge cases (empty sites, one
site, tiny sites in slices, a site larger than the shared-memory q table).  --time: device-resident passes by CUDA events.
When a bounded wait of the kernel traps, the trap record (wait site, block, thread, parity) is printed."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from m6anet_b200 import _cabi, weights as W                       # noqa: E402
from m6anet_b200.engine import MilEngine                          # noqa: E402
from conftest import ALL_TAGS, GOLDEN, oracle_params              # noqa: E402
from oracle import mil_inference, read_probabilities_float64      # noqa: E402  (checker)
from oracle import c_oracle                                       # noqa: E402  (checker)

NPZ = {"HCT116_RNA002": "rna002_hct116.npz", "arabidopsis_RNA002": "rna002_arabidopsis_virc.npz",
       "HEK293T_RNA004": "rna004_hek293t_glori.npz", "HEK293T_RNA004_M6ACE": "rna004_hek293t_m6ace.npz"}
WAIT_SITES = {1: "x_full (MMA)", 2: "l1_done (E1)", 3: "a_full (MMA)", 4: "a_free (E1)", 5: "d2_full (E2)",
              6: "d2_free (MMA)", 7: "slab_full (MC)", 8: "slab_empty (staging)", 9: "x_full (E1)", 10: "l1_done (staging)", 11: "hdr_ready (staging)"}
res = {"cases": [], "ok": True}


def engine_for(tag):
    if tag in NPZ:
        return MilEngine(W.from_npz(os.path.join(ROOT, "m6anet_b200", "assets", "model_states", NPZ[tag])), 0)
    from conftest import load_golden
    g = load_golden(tag)
    return MilEngine(W.from_raw({k[2:]: g[k] for k in g if k.startswith("w_")}), 0)


def run(eng, feats, off, kmer, n_iters, **kw):
    try:
        return eng.infer_host(feats, off, kmer, n_iters, **kw)
    except Exception as e:                                           # noqa: BLE001
        print(f"KERNEL FAILED: {e}", flush=True)
        for rec in eng.trap_record():
            print(f"   stuck wait: site {rec[0]} = {WAIT_SITES.get(rec[0], '?')}, block {rec[1]}, thread {rec[2]} "
                  f"(warp {rec[2] // 32}), parity {rec[3]}", flush=True)
        raise


def check(name, eng, P, feats, off, kmer, n_iters, seed=0, base=0, thr=0.033379376, read_bar=2e-6):
    eng.set_encoder("tc")
    rp, sp, mc = run(eng, feats, off, kmer, n_iters, seed=seed, site_id_base=base, read_threshold=thr)
    eng.set_encoder("ffma")
    rp_f, sp_f, mc_f = run(eng, feats, off, kmer, n_iters, seed=seed, site_id_base=base, read_threshold=thr)
    n_reads = np.diff(off)
    rows = None if P.emb is None else np.repeat(kmer, n_reads, axis=0)
    orp, osp, omc = (c_oracle.mil_inference if len(off) > 2000 else mil_inference)(
        P, feats, off, kmer, n_iters=n_iters, seed=seed, site_id_base=base, read_threshold=thr)
    p64 = read_probabilities_float64(P, feats, rows) if len(feats) else np.zeros(0)
    live = n_reads > 0
    d = {
        "case": name, "sites": int(len(off) - 1), "reads": int(len(feats)),
        "read_vs_p32": float(np.abs(rp - orp).max()) if len(feats) else 0.0,
        "read_vs_p64": float(np.abs(rp - p64).max()) if len(feats) else 0.0,
        "ffma_read_vs_p64": float(np.abs(rp_f - p64).max()) if len(feats) else 0.0,
        "site_vs_oracle": float(np.abs(sp[live] - osp[live]).max()) if live.any() else 0.0,
        "site_tc_vs_ffma": float(np.abs(sp[live] - sp_f[live]).max()) if live.any() else 0.0,
        "nan_sites_ok": bool(np.all(np.isnan(sp[~live])) and np.all(mc[~live] == 0)),
        "mod_count_diffs": int((mc != omc).sum()), "mod_count_tc_vs_ffma": int((mc != mc_f).sum()),
    }
    ok = d["read_vs_p32"] <= read_bar and d["read_vs_p64"] <= 3e-6 and d["site_vs_oracle"] <= 1e-4 and d["nan_sites_ok"]
    d["ok"] = bool(ok)
    res["cases"].append(d)
    res["ok"] = res["ok"] and ok
    print(("ok   " if ok else "FAIL ") + json.dumps(d), flush=True)
    return rp, sp, mc


z = np.load(os.path.join(GOLDEN, "synthetic_inputs.npz"))
feats, off, kmer = z["feats"], z["read_off"], z["kmer_idx"]
print(_cabi.lib().m6a_build_info().decode(), flush=True)
for tag in ([] if "--no-parity" in sys.argv else (ALL_TAGS[:1] if "--first-only" in sys.argv else ALL_TAGS)):
    eng = engine_for(tag)
    eng.trap_record()            # arm
    P = oracle_params(tag)
    thr = 0.0032978046219796 if tag.startswith("arabidopsis") else 0.033379376
    check(f"golden/{tag}", eng, P, feats, off, kmer, 200, seed=1234, base=7_000_000_000, thr=thr)
    if tag == "HCT116_RNA002" and "--first-only" not in sys.argv:
        rng = np.random.default_rng(7)
        # one site / a partial tile
        check("one site x 20", eng, P, feats[:20], off[:2], kmer[:1], 1000)
        # empty sites in between
        n = np.array([0, 25, 0, 0, 40, 130, 0], dtype=np.int64)
        o = np.concatenate([[0], np.cumsum(n)])
        f = rng.standard_normal((int(o[-1]), 9), dtype=np.float32)
        check("empty sites", eng, P, f, o, kmer[:7], 300, seed=3)
        # tiny sites: more than 64 sites per tile -> slices
        n = rng.integers(1, 6, size=3000).astype(np.int64)
        o = np.concatenate([[0], np.cumsum(n)])
        f = rng.standard_normal((int(o[-1]), 9), dtype=np.float32)
        k = rng.integers(0, 66, size=(3000, 3)).astype(np.int32)
        check("3000 tiny sites", eng, P, f, o, k, 64, seed=5)
        # a site larger than the q table, between ordinary ones
        n = np.array([30, 5000, 22, 4096, 4097, 50], dtype=np.int64)
        o = np.concatenate([[0], np.cumsum(n)])
        f = rng.standard_normal((int(o[-1]), 9), dtype=np.float32)
        check("huge sites", eng, P, f, o, kmer[:6], 100, seed=9)
        # short uniform jobs: the tile size is balanced over the SMs (tiles of 3-4 MMA tiles, one tile per CTA)
        for S_ in (300, 1000, 5000):
            f = rng.standard_normal((S_ * 50, 9), dtype=np.float32)
            o = np.arange(S_ + 1, dtype=np.int64) * 50
            k = rng.integers(0, 66, size=(S_, 3)).astype(np.int32)
            check(f"{S_} x 50", eng, P, f, o, k, 1000, seed=2)
        # a uniform job large enough for every CTA: 40 000 sites x 50 reads, every site against the C oracle
        S, nr_ = 40_000, 50
        f = rng.standard_normal((S * nr_, 9), dtype=np.float32)
        o = np.arange(S + 1, dtype=np.int64) * nr_
        k = rng.integers(0, 66, size=(S, 3)).astype(np.int32)
        check("40k x 50", eng, P, f, o, k, 1000, seed=0, read_bar=5e-6)
        # ragged lognormal
        n = np.clip(np.round(np.exp(rng.normal(np.log(33), 0.8, size=20000))), 20, 1000).astype(np.int64)
        o = np.concatenate([[0], np.cumsum(n)])
        f = rng.standard_normal((int(o[-1]), 9), dtype=np.float32)
        k = rng.integers(0, 66, size=(20000, 3)).astype(np.int32)
        check("20k ragged", eng, P, f, o, k, 1000, seed=11, read_bar=5e-6)
    eng.close()

if "--time" in sys.argv:
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rt.cudaGetErrorString.restype = C.c_char_p

    def dev(a=None, nbytes=0):
        p = C.c_void_p()
        n = a.nbytes if a is not None else nbytes
        assert rt.cudaMalloc(C.byref(p), max(n, 16)) == 0
        if a is not None and n:
            assert rt.cudaMemcpy(p, a.ctypes.data_as(C.c_void_p), n, 1) == 0
        return p

    rng = np.random.default_rng(0)
    sizes = [(200_000, 50), (125_000, 50)]
    if "--big" in sys.argv:
        sizes.append((1_000_000, 50))
    if "--only-big" in sys.argv:
        sizes = [(1_000_000, 50)]
    for arg in sys.argv:
        if arg.startswith("--sizes="):          # --sizes=1000000:32,1000000:64
            sizes = [tuple(int(v) for v in x.split(":")) for x in arg.split("=", 1)[1].split(",")]
    encs = ("tc",) if "--only-tc" in sys.argv else ("tc", "ffma")
    eng = engine_for("HCT116_RNA002")
    eng.trap_record()
    PL = _cabi.lib()
    ev = [C.c_void_p(), C.c_void_p()]
    for e in ev:
        assert rt.cudaEventCreate(C.byref(e)) == 0
    ms = C.c_float()
    res["timing"] = []
    for S, n in sizes:
        R = S * n
        tf = rng.standard_normal((R, 9), dtype=np.float32)
        to = np.arange(S + 1, dtype=np.int64) * n
        tk = rng.integers(0, 66, size=(S, 3)).astype(np.int32)
        ws = int(PL.m6a_mil_workspace_bytes(R))
        d = [dev(tf), dev(to), dev(tk), dev(nbytes=4 * R), dev(nbytes=4 * S), dev(nbytes=4 * S), dev(nbytes=ws)]
        for enc in encs:
            eng.set_encoder(enc)
            for iters in (1000, 1):
                best = 1e9
                for it in range(5):
                    rt.cudaEventRecord(ev[0], None)
                    rc = PL.m6a_mil_infer_f32(eng._handle, d[0], d[1], d[2], S, R, 0, 20, iters, 0, None, 0.033379376, d[3], d[4],
                                              d[5], d[6], ws, None)
                    assert rc == 0, rc
                    rt.cudaEventRecord(ev[1], None)
                    e = rt.cudaEventSynchronize(ev[1])
                    if e != 0:
                        print(f"timing run failed: {rt.cudaGetErrorString(e)}; stuck waits {eng.trap_record()}")
                        sys.exit(3)
                    rt.cudaEventElapsedTime(C.byref(ms), ev[0], ev[1])
                    if it:
                        best = min(best, ms.value)
                row = {"sites": S, "reads": n, "encoder": enc, "iters": iters, "ms": best, "launch": eng.last_launch()}
                res["timing"].append(row)
                print(json.dumps(row), flush=True)
        for p in d:
            rt.cudaFree(p)

out_path = next((x for x in sys.argv[1:] if not x.startswith("--")), None)
if out_path:
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump(res, fh, indent=1)
print("OK" if res["ok"] else "FAIL")
sys.exit(0 if res["ok"] else 1)
