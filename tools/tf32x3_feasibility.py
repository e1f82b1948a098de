#!/usr/bin/env python
"""Numerics feasibility of a tensor-core read encoder (NOT built): would a tcgen05 kind::tf32 encoder
with the usual error-compensated operand split keep the parity bars?

Emulates on the CPU, for the golden inputs, Linear-1 and Linear-2 evaluated as sums of TF32 x TF32 products with float32
accumulation, for the split variants
    1xTF32   a_hi*b_hi
    2xTF32   a_hi*b_hi + a_lo*b_hi                (activations split, weights rounded once)
    3xTF32   a_hi*b_hi + a_lo*b_hi + a_hi*b_lo    (both split; the a_lo*b_lo term is dropped)
and reports max |p - p_float64| per read next to the float32 CUDA-core formulation (the oracle).
TF32 rounding = round-to-nearest-even to 10 explicit mantissa bits.  Accumulation inside the tensor core is modelled as
float32 adds of exact products (an optimistic model: the hardware may truncate; the GPU run decides).  The last line is
the pessimistic model of the kernel as written: operands truncated to TF32, and the accumulator rounded TOWARD ZERO to
float32 after every K = 8 step of every term (20 steps x 3 terms for Linear-2)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ReadEncoderParams, read_probabilities   # noqa: E402


def tf32(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    lsb = (u >> np.uint64(13)) & np.uint64(1)
    u = (u + np.uint64(0xFFF) + lsb) & np.uint64(0xFFFFE000)
    return u.astype(np.uint32).view(np.float32).reshape(x.shape)


def mm(a, b, terms):
    """a [N,K] @ b [K,M] with TF32 operands, products exact in float64, accumulated to float32 once per term."""
    a_hi, b_hi = tf32(a), tf32(b)
    a_lo, b_lo = tf32(a - a_hi), tf32(b - b_hi)
    acc = (a_hi.astype(np.float64) @ b_hi.astype(np.float64)).astype(np.float32)
    if terms >= 2:
        acc = acc + (a_lo.astype(np.float64) @ b_hi.astype(np.float64)).astype(np.float32)
    if terms >= 3:
        acc = acc + (a_hi.astype(np.float64) @ b_lo.astype(np.float64)).astype(np.float32)
    return acc


def main():
    P = ReadEncoderParams.from_npz(os.path.join(ROOT, "m6anet_b200", "assets", "model_states", "rna002_hct116.npz"))
    z = np.load(os.path.join(ROOT, "tests", "golden", "synthetic_inputs.npz"))
    feats, off, kmer = z["feats"], z["read_off"], z["kmer_idx"]
    rows = np.repeat(kmer, np.diff(off), axis=0)
    s = (P.bn_gamma.astype(np.float64) / np.sqrt(P.bn_var.astype(np.float64) + P.bn_eps))
    w1 = (P.w1.astype(np.float64) * s[:, None])
    b1 = ((P.b1.astype(np.float64) - P.bn_mean) * s + P.bn_beta)
    x = np.concatenate([feats, P.emb[rows].reshape(-1, 6)], axis=1).astype(np.float64)
    # float64 truth
    h = np.maximum(x @ w1.T + b1, 0)
    h2 = np.maximum(h @ P.w2.astype(np.float64).T + P.b2, 0)
    p64 = 1 / (1 + np.exp(-(h2 @ P.w3.astype(np.float64).reshape(-1) + float(P.b3[0]))))
    p32 = read_probabilities(P, feats, rows)
    print(f"{len(p64)} reads; float32 CUDA-core formulation (oracle): max|p - p64| = {np.abs(p32 - p64).max():.2e}")
    w1f, b1f = w1.astype(np.float32), b1.astype(np.float32)
    for terms in (1, 2, 3):
        hh = np.maximum(mm(x.astype(np.float32), w1f.T.copy(), terms) + b1f, 0)
        hh2 = np.maximum(mm(hh, P.w2.T.copy(), terms) + P.b2, 0)
        zz = hh2.astype(np.float32) @ P.w3.reshape(-1) + P.b3[0]
        p = (1 / (1 + np.exp(-zz.astype(np.float32)))).astype(np.float32)
        print(f"{terms}xTF32: max|p - p64| = {np.abs(p - p64).max():.2e}   max|p - p_oracle32| = {np.abs(p - p32).max():.2e}")


    # pessimistic model: truncation everywhere, round-toward-zero accumulation per MMA (K = 8) step
    def tr(v):
        v = np.ascontiguousarray(v, dtype=np.float32)
        return (v.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32).reshape(v.shape)

    def rz32(v64):
        f = v64.astype(np.float32)
        f = np.where(np.abs(f.astype(np.float64)) > np.abs(v64), np.nextafter(f, np.float32(0)), f)
        return f.astype(np.float32)

    def mm_rz(a, b):
        ah, bh = tr(a), tr(b)
        al, bl = tr(a - ah), tr(b - bh)
        acc = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
        for k0 in range(0, a.shape[1], 8):
            sl = slice(k0, k0 + 8)
            for u, v in ((ah, bh), (al, bh), (ah, bl)):
                acc = rz32(acc.astype(np.float64) + u[:, sl].astype(np.float64) @ v[:, sl].astype(np.float64).T)
        return acc

    xa = np.zeros((len(x), 16), dtype=np.float32)
    xa[:, :15] = x.astype(np.float32)
    xa[:, 15] = 1.0
    w1a = np.zeros((160, 16), dtype=np.float32)
    w1a[:150, :15] = w1f
    w1a[:150, 15] = b1f
    w2a = np.zeros((32, 160), dtype=np.float32)
    w2a[:, :150] = P.w2
    hh2 = mm_rz(np.maximum(mm_rz(xa, w1a), 0), w2a)
    zz = np.maximum(hh2 + P.b2, 0) @ P.w3.reshape(-1) + P.b3[0]
    p = (1 / (1 + np.exp(-zz.astype(np.float32)))).astype(np.float32)
    print(f"3xTF32, truncated operands, accumulator rounded toward zero per K-step: max|p - p64| = {np.abs(p - p64).max():.2e}")


if __name__ == "__main__":
    main()
