#!/usr/bin/env python
"""CPU model of the tensor-core read encoder's numerics (m6anet_b200/csrc/m6a_kernel_tc.cu) -- a design tool, not a test.

Emulates Linear-1 / Linear-2 as sums of TF32 x TF32 products for the golden inputs and every weight set, under the behaviour
MEASURED on the B200 (tools/microbench/tcgen05_probe.cu: 32-bit operands are truncated to TF32) and the accumulation model
that reproduces the measured errors of round 1's single-accumulator kernel to ~10 %: the float32 accumulator is rounded
TOWARD ZERO after every MMA (K = 8) step.  Variants: how the operands are split (truncate / round-to-nearest hi, lo), whether
the two small terms get their own accumulator (split_acc), and whether accumulators are read out every kchunk_acc K-steps.
Prints max |p - p_float64| and max |p - p_float32 oracle| per variant.  The shipped kernel is "M5": rn-hi / trunc-lo with
split accumulators in Linear-2.      python tools/tc_numerics_model.py"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+"/tests")
from oracle import ReadEncoderParams, read_probabilities, read_probabilities_float64
from conftest import ALL_TAGS, oracle_params
z = np.load(ROOT+"/tests/golden/synthetic_inputs.npz")
feats, off, kmer = z["feats"], z["read_off"], z["kmer_idx"]
rows = np.repeat(kmer, np.diff(off), axis=0)
def tr(v):
    v = np.ascontiguousarray(v, dtype=np.float32); return (v.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32).reshape(v.shape)
def rn(v):   # round half away (bits + 0x1000) & mask
    v = np.ascontiguousarray(v, dtype=np.float32); return ((v.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32).reshape(v.shape)
def rz32(v64):
    f = v64.astype(np.float32)
    f = np.where(np.abs(f.astype(np.float64)) > np.abs(v64), np.nextafter(f, np.float32(0)), f)
    return f.astype(np.float32)
def rn32(v64): return v64.astype(np.float32)
def mm(a, b, hi_a, lo_a, hi_b, lo_b, acc_round, split_acc, kchunk_acc):
    ah, bh = hi_a(a), hi_b(b); al, bl = lo_a(a - ah), lo_b(b - bh)
    N, K = a.shape; M = b.shape[0]
    main = np.zeros((N, M), np.float32); corr = np.zeros((N, M), np.float32); total = np.zeros((N,M), np.float32)
    for k0 in range(0, K, 8):
        sl = slice(k0, k0+8)
        prods = [(ah, bh, 0), (al, bh, 1), (ah, bl, 1)]
        for u, v, is_corr in prods:
            pr = u[:, sl].astype(np.float64) @ v[:, sl].astype(np.float64).T
            if split_acc and is_corr: corr = acc_round(corr.astype(np.float64) + pr)
            else: main = acc_round(main.astype(np.float64) + pr)
        if kchunk_acc and ((k0 // 8) % kchunk_acc == kchunk_acc - 1):
            total = (total + main) + corr if split_acc else total + main   # fp32 RN adds on CUDA cores
            main[:] = 0; corr[:] = 0
    if kchunk_acc: return (total + main) + corr if split_acc else total + main
    return main + corr if split_acc else main
def run(P, variant):
    s = (P.bn_gamma.astype(np.float64) / np.sqrt(P.bn_var.astype(np.float64) + P.bn_eps))
    w1 = (P.w1.astype(np.float64) * s[:, None]); b1 = ((P.b1.astype(np.float64) - P.bn_mean) * s + P.bn_beta)
    E = 0 if P.emb is None else P.emb.shape[1]
    x = feats if E == 0 else np.concatenate([feats, P.emb[rows].reshape(-1, 3*E)], axis=1)
    xa = np.zeros((len(x), 16), np.float32); xa[:, :x.shape[1]] = x; xa[:, 15] = 1.0
    w1a = np.zeros((160, 16), np.float32); w1a[:150, :x.shape[1]] = w1.astype(np.float32); w1a[:150, 15] = b1.astype(np.float32)
    w2a = np.zeros((32, 160), np.float32); w2a[:, :150] = P.w2
    h = np.maximum(mm(xa, w1a, **variant), 0)
    v2 = dict(variant)
    h2 = mm(h, w2a, **v2)
    zz = (np.maximum(h2 + P.b2, 0) @ P.w3.reshape(-1) + P.b3[0]).astype(np.float32)
    return (1 / (1 + np.exp(-zz))).astype(np.float32)
V = {
 "M0 trunc/trunc RZacc":            dict(hi_a=tr, lo_a=tr, hi_b=tr, lo_b=tr, acc_round=rz32, split_acc=False, kchunk_acc=0),
 "M0r trunc/trunc RNacc":           dict(hi_a=tr, lo_a=tr, hi_b=tr, lo_b=tr, acc_round=rn32, split_acc=False, kchunk_acc=0),
 "M1 rn-hi/trunc-lo RZacc":         dict(hi_a=rn, lo_a=tr, hi_b=rn, lo_b=rn, acc_round=rz32, split_acc=False, kchunk_acc=0),
 "M2 rn/rn RZacc":                  dict(hi_a=rn, lo_a=rn, hi_b=rn, lo_b=rn, acc_round=rz32, split_acc=False, kchunk_acc=0),
 "M3 rn/rn RZacc splitacc":         dict(hi_a=rn, lo_a=rn, hi_b=rn, lo_b=rn, acc_round=rz32, split_acc=True, kchunk_acc=0),
 "M4 rn/rn RZacc splitacc chunk4":  dict(hi_a=rn, lo_a=rn, hi_b=rn, lo_b=rn, acc_round=rz32, split_acc=True, kchunk_acc=4),
 "M5 rn-hi/trunc-lo RZ splitacc":   dict(hi_a=rn, lo_a=tr, hi_b=rn, lo_b=rn, acc_round=rz32, split_acc=True, kchunk_acc=0),
 "M6 trunc/trunc RZ splitacc":      dict(hi_a=tr, lo_a=tr, hi_b=tr, lo_b=tr, acc_round=rz32, split_acc=True, kchunk_acc=0),
}
for tag in ALL_TAGS:
    P = oracle_params(tag)
    p64 = read_probabilities_float64(P, feats, None if P.emb is None else rows)
    p32 = read_probabilities(P, feats, None if P.emb is None else rows)
    print(f"{tag}: oracle32 vs p64 {np.abs(p32-p64).max():.2e}")
    for name, v in V.items():
        p = run(P, v)
        print(f"   {name:34s} |p-p64| {np.abs(p-p64).max():.2e}  |p-p32| {np.abs(p-p32).max():.2e}")
