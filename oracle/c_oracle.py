"""ctypes wrapper of the C restatement (oracle/c/m6a_oracle.c) -- TEST INFRASTRUCTURE ONLY.
Used where the NumPy oracle would be too slow: checking EVERY site of a full-size job."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .mil_oracle import ReadEncoderParams

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_DIR, "libm6a_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", _DIR], check=True, capture_output=True)
        _LIB = C.CDLL(path)
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def sample_indices(seed, site_id, n_reads, n_iters, n_samples=20):
    out = np.empty((n_iters, n_samples), dtype=np.int32)
    lib().oracle_sample_indices(C.c_uint64(seed & (2**64 - 1)), C.c_uint64(site_id & (2**64 - 1)), C.c_uint32(n_reads),
                                C.c_int(n_iters), C.c_int(n_samples), _p(out))
    return out


def mil_inference(params: ReadEncoderParams, feats, read_off, kmer_idx, n_iters, seed=0, site_id_base=0, n_samples=20,
                  read_threshold=0.033379376):
    """Same contract as oracle.mil_inference, multi-threaded C."""
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    feats = f32(feats)
    read_off = np.ascontiguousarray(read_off, dtype=np.int64)
    n_sites = len(read_off) - 1
    emb = None if params.emb is None else f32(params.emb)
    kid = np.zeros((n_sites, 3), np.int32) if kmer_idx is None else np.ascontiguousarray(kmer_idx, dtype=np.int32)
    rp = np.empty(feats.shape[0], dtype=np.float32)
    sp = np.empty(n_sites, dtype=np.float32)
    mc = np.empty(n_sites, dtype=np.int32)
    keep = [f32(a) for a in (params.w1, params.b1, params.bn_gamma, params.bn_beta, params.bn_mean, params.bn_var, params.w2,
                            params.b2, params.w3, params.b3)]
    L = lib()
    n_thr = max(1, min(len(os.sched_getaffinity(0)), n_sites))
    cuts = [n_sites * t // n_thr for t in range(n_thr + 1)]

    def work(t):   # ctypes releases the GIL: site ranges run on all host cores
        lo, hi = cuts[t], cuts[t + 1]
        L.oracle_read_probs(_p(feats), _p(read_off), _p(kid), C.c_int64(lo), C.c_int64(hi), C.c_int(params.w1.shape[0]),
                            C.c_int(params.w2.shape[0]), C.c_int(0 if emb is None else emb.shape[1]), _p(emb),
                            _p(keep[0]), _p(keep[1]), _p(keep[2]), _p(keep[3]), _p(keep[4]), _p(keep[5]), C.c_float(params.bn_eps),
                            _p(keep[6]), _p(keep[7]), _p(keep[8]), _p(keep[9]), _p(rp))
        L.oracle_site_probs(_p(rp), _p(read_off), C.c_int64(lo), C.c_int64(hi), C.c_int64(site_id_base), C.c_int(n_iters),
                            C.c_int(n_samples), C.c_uint64(seed & (2**64 - 1)), C.c_float(read_threshold), _p(sp), _p(mc))

    with ThreadPoolExecutor(n_thr) as ex:
        list(ex.map(work, range(n_thr)))
    return rp, sp, mc


def sample_bags(seed, site_id, n_reads, n_iters, n_samples=20):
    out = np.empty((n_iters, n_samples), dtype=np.int32)
    lib().oracle_sample_bags(C.c_uint64(seed & (2**64 - 1)), C.c_uint64(site_id & (2**64 - 1)), C.c_uint32(n_reads),
                             C.c_int(n_iters), C.c_int(n_samples), _p(out))
    return out


def mil_validate(params: ReadEncoderParams, feats, read_off, kmer_idx, n_iters, seed=0, site_id_base=0, n_samples=20,
                 pool="prod"):
    """(read_prob, bag_prob [n_sites, n_iters]) like oracle.mil_validate on the device bag stream, multi-threaded C."""
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    feats = f32(feats)
    read_off = np.ascontiguousarray(read_off, dtype=np.int64)
    n_sites = len(read_off) - 1
    emb = None if params.emb is None else f32(params.emb)
    kid = np.zeros((n_sites, 3), np.int32) if kmer_idx is None else np.ascontiguousarray(kmer_idx, dtype=np.int32)
    rp = np.empty(feats.shape[0], dtype=np.float32)
    bag = np.empty((n_sites, n_iters), dtype=np.float32)
    keep = [f32(a) for a in (params.w1, params.b1, params.bn_gamma, params.bn_beta, params.bn_mean, params.bn_var, params.w2,
                            params.b2, params.w3, params.b3)]
    L = lib()
    n_thr = max(1, min(len(os.sched_getaffinity(0)), n_sites))
    cuts = [n_sites * t // n_thr for t in range(n_thr + 1)]
    code = {"prod": 0, "mean": 1, "max": 2}[pool]

    def work(t):
        lo, hi = cuts[t], cuts[t + 1]
        L.oracle_read_probs(_p(feats), _p(read_off), _p(kid), C.c_int64(lo), C.c_int64(hi), C.c_int(params.w1.shape[0]),
                            C.c_int(params.w2.shape[0]), C.c_int(0 if emb is None else emb.shape[1]), _p(emb),
                            _p(keep[0]), _p(keep[1]), _p(keep[2]), _p(keep[3]), _p(keep[4]), _p(keep[5]), C.c_float(params.bn_eps),
                            _p(keep[6]), _p(keep[7]), _p(keep[8]), _p(keep[9]), _p(rp))
        L.oracle_bag_probs(_p(rp), _p(read_off), C.c_int64(lo), C.c_int64(hi), C.c_int64(site_id_base), C.c_int(n_iters),
                           C.c_int(n_samples), C.c_uint64(seed & (2**64 - 1)), C.c_int(code), _p(bag))

    with ThreadPoolExecutor(n_thr) as ex:
        list(ex.map(work, range(n_thr)))
    return rp, bag
