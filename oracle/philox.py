"""Index streams shared by the CUDA kernel and the oracle (test infrastructure).

The reference draws its bags with ``np.random.choice(proba, n_iters * n_samples,
replace=True)`` (reference m6anet/utils/inference_utils.py:85), i.e. iid uniform indices
from the process-global MT19937 stream, inside a freshly forked ``Pool`` (ibid. :103), so
the reference's own output is not reproducible run to run.  Parity on
``probability_modified`` is therefore defined on a SHARED index stream.  Two streams exist:

* ``sample_indices`` -- the product's device stream: Philox4x32-10 (Salmon et al., SC'11;
  the generator behind cuRAND/PyTorch CUDA), counter-based so that the draw for
  (seed, global site id, iteration, sample) does not depend on GPU count or tiling.
  This file is its NumPy restatement; ``m6anet_b200/csrc/m6a_philox.cuh`` is the device copy.
* ``sample_indices_mt19937`` -- a replay of the reference's legacy ``np.random`` stream for
  ONE site drawn right after ``np.random.seed`` (usable through the kernel's explicit-index
  mode).
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)
_SHIFT32 = np.uint64(32)


def philox4x32_10(ctr, key):
    """Philox4x32 with 10 rounds, vectorised.

    ctr: uint32 array [..., 4]; key: uint32 array broadcastable to [..., 2].
    Returns uint32 array [..., 4].
    """
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    c0 = ctr[..., 0].astype(np.uint64)
    c1 = ctr[..., 1].astype(np.uint64)
    c2 = ctr[..., 2].astype(np.uint64)
    c3 = ctr[..., 3].astype(np.uint64)
    k0 = np.broadcast_to(key[..., 0], c0.shape).astype(np.uint64)
    k1 = np.broadcast_to(key[..., 1], c0.shape).astype(np.uint64)
    for r in range(10):
        if r:
            k0 = (k0 + np.uint64(PHILOX_W0)) & _MASK32
            k1 = (k1 + np.uint64(PHILOX_W1)) & _MASK32
        p0 = PHILOX_M0 * c0          # 32x32 -> 64, no overflow in uint64
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> _SHIFT32, p0 & _MASK32
        hi1, lo1 = p1 >> _SHIFT32, p1 & _MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
    out = np.stack([c0, c1, c2, c3], axis=-1)
    return out.astype(np.uint32)


def sample_indices(seed: int, site_id: int, n_reads: int, n_iters: int, n_samples: int = 20):
    """Device index stream for one site -> int64 [n_iters, n_samples] in [0, n_reads).

    Specification (mirrored by the CUDA kernel):
      key   = (seed & 0xffffffff, seed >> 32)
      ctr   = (call, iteration, site_id & 0xffffffff, site_id >> 32),  call = sample // 4
      word  = philox4x32_10(ctr, key)[sample % 4]
      index = (word * n_reads) >> 32            (multiply-shift; bias <= n_reads / 2**32)
    """
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    site_id = int(site_id) & 0xFFFFFFFFFFFFFFFF
    n_calls = (n_samples + 3) // 4
    ctr = np.empty((n_iters, n_calls, 4), dtype=np.uint32)
    ctr[..., 0] = np.arange(n_calls, dtype=np.uint32)[None, :]
    ctr[..., 1] = np.arange(n_iters, dtype=np.uint32)[:, None]
    ctr[..., 2] = site_id & 0xFFFFFFFF
    ctr[..., 3] = site_id >> 32
    key = np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32)
    words = philox4x32_10(ctr, key).reshape(n_iters, n_calls * 4)[:, :n_samples]
    return ((words.astype(np.uint64) * np.uint64(n_reads)) >> _SHIFT32).astype(np.int64)


def sample_indices_mt19937(seed: int, n_reads: int, n_iters: int, n_samples: int = 20):
    """Indices the reference's ``np.random.choice(proba, n_iters*n_samples, replace=True)``
    uses when called first after ``np.random.seed(seed)`` (reference
    m6anet/utils/inference_utils.py:85, seeding at m6anet/scripts/inference.py:86).
    Legacy ``RandomState.choice`` with replacement draws ``randint(0, n, size)``.
    """
    rs = np.random.RandomState(seed)
    return rs.randint(0, n_reads, size=n_iters * n_samples).reshape(n_iters, n_samples).astype(np.int64)
