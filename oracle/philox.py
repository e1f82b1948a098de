"""Index streams shared by the CUDA kernel and the oracle (test infrastructure).

The reference draws its bags with ``np.random.choice(proba, n_iters * n_samples,
replace=True)`` (reference m6anet/utils/inference_utils.py:85), i.e. iid uniform indices
from the process-global MT19937 stream, inside a freshly forked ``Pool`` (ibid. :103), so
the reference's own output is not reproducible run to run.  Parity on
``probability_modified`` is therefore defined on a SHARED index stream.  Two streams exist:

* ``sample_indices`` -- the product's device stream (specification below); this file is its
  NumPy restatement, ``m6anet_b200/csrc/m6a_rng.cuh`` the device copy; the two are compared
  bit for bit on the GPU (tests/test_gpu_parity.py::test_device_index_stream_matches_oracle).
* ``sample_indices_mt19937`` -- a replay of the reference's legacy ``np.random`` stream for
  ONE site drawn right after ``np.random.seed`` (usable through the kernel's explicit-index
  mode).

Device stream specification ("Philox-seeded MWC64X lane streams")
-----------------------------------------------------------------
A site's iterations are cut into blocks of ``32 * ipl`` iterations,
``ipl = max(8, ceil(n_iters / 2048))`` (so at most 64 blocks).  Iteration ``it`` belongs to
block ``b = it // (32*ipl)``, lane ``l = it % 32`` and is that lane's round ``k = (it // 32) % ipl``.
Every (site, block, lane) owns an independent generator:

  seeding   (w0, w1, _, _) = Philox4x32-10(ctr = (l, b, site_id & 0xffffffff, site_id >> 32),
                                           key = (seed & 0xffffffff, seed >> 32))
            x = w0;  c = (w1 * (A - 1)) >> 32;  if x == c == 0: x = 1          (A = 4294883355)
  draw      word = x ^ c;  t = A * x + c (64 bit);  x = t & 0xffffffff;  c = t >> 32     (MWC64X, D. B. Thomas 2011)
  index     n_reads >  256: one index per word:   (word * n_reads) >> 32               (bias <= n_reads / 2**32)
            n_reads <= 256: two indices per word: u = word * n_reads (64 bit);  i1 = u >> 32;
                            i2 = ((u & 0xffffffff) * n_reads) >> 32                    (bias <= n_reads**2 / 2**32 <= 1.6e-5)
            (the low half of the first product is again uniform on a lattice of spacing n_reads; the paired regime
             halves the generator steps and the quarter-rate high multiplies on the device).  An iteration consumes
             ceil(n_samples / 2) words in the paired regime; an unused odd half is dropped.

and draws, in order, the ``n_samples`` indices of its round 0, then round 1, ... .  Philox4x32-10
(Salmon et al., SC'11; the cuRAND / PyTorch-CUDA generator) gives key/counter separation, so a
site's stream does not depend on GPU count, sharding or tiling; the multiply-with-carry stream
costs one wide multiply per draw on the device instead of twenty.
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MWC_A = 4294883355
_MASK32 = np.uint64(0xFFFFFFFF)
_SHIFT32 = np.uint64(32)
MAX_BLOCKS = 64
MIN_ITERS_PER_LANE = 8
PAIRED_MAX_READS = 256


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


def iters_per_lane(n_iters: int) -> int:
    return max(MIN_ITERS_PER_LANE, -(-int(n_iters) // (32 * MAX_BLOCKS)))


def block_layout(n_iters: int):
    """(ipl, n_blocks) of the device decomposition of a site's iterations."""
    ipl = iters_per_lane(n_iters)
    return ipl, -(-int(n_iters) // (32 * ipl))


def _streams(seed: int, site_ids: np.ndarray, n_blocks: int):
    """Philox-seeded MWC64X states (x, c), uint64 arrays [S, n_blocks, 32]."""
    S = len(site_ids)
    ctr = np.empty((S, n_blocks, 32, 4), dtype=np.uint32)
    ctr[..., 0] = np.arange(32, dtype=np.uint32)[None, None, :]
    ctr[..., 1] = np.arange(n_blocks, dtype=np.uint32)[None, :, None]
    ctr[..., 2] = (site_ids & _MASK32).astype(np.uint32)[:, None, None]
    ctr[..., 3] = (site_ids >> _SHIFT32).astype(np.uint32)[:, None, None]
    key = np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32)
    w = philox4x32_10(ctr, key)
    x = w[..., 0].astype(np.uint64)
    c = (w[..., 1].astype(np.uint64) * np.uint64(MWC_A - 1)) >> _SHIFT32
    x = np.where((x == 0) & (c == 0), np.uint64(1), x)
    return x, c


def _draw_group(seed, site_ids, n_reads, n_iters, n_samples, paired: bool) -> np.ndarray:
    """All sites of one regime (single / paired) -> int64 [S, n_iters, n_samples]."""
    S = len(site_ids)
    ipl, n_blocks = block_layout(n_iters)
    x, c = _streams(seed, site_ids, n_blocks)
    A = np.uint64(MWC_A)
    out = np.empty((S, n_blocks, ipl, 32, n_samples), dtype=np.int64)   # [S, block, round, lane, sample]
    nr = n_reads[:, None, None]
    for k in range(ipl):
        s = 0
        while s < n_samples:
            word = x ^ c
            t = A * x + c
            x, c = t & _MASK32, t >> _SHIFT32
            u = word * nr                                   # < 2**64
            out[:, :, k, :, s] = (u >> _SHIFT32).astype(np.int64)
            s += 1
            if paired and s < n_samples:
                out[:, :, k, :, s] = (((u & _MASK32) * nr) >> _SHIFT32).astype(np.int64)
                s += 1
    return out.reshape(S, n_blocks * ipl * 32, n_samples)[:, :n_iters, :]


def sample_indices_many(seed: int, site_ids, n_reads, n_iters: int, n_samples: int = 20) -> np.ndarray:
    """Device index stream for several sites -> int64 [n_sites, n_iters, n_samples]."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    site_ids = np.asarray(site_ids, dtype=np.uint64).reshape(-1)
    n_reads = np.asarray(n_reads, dtype=np.uint64).reshape(-1)
    out = np.empty((len(site_ids), n_iters, n_samples), dtype=np.int64)
    paired = n_reads <= PAIRED_MAX_READS
    for flag in (False, True):
        sel = np.nonzero(paired == flag)[0]
        if len(sel):
            out[sel] = _draw_group(seed, site_ids[sel], n_reads[sel], n_iters, n_samples, flag)
    return out


def sample_indices(seed: int, site_id: int, n_reads: int, n_iters: int, n_samples: int = 20) -> np.ndarray:
    """Device index stream for one site -> int64 [n_iters, n_samples] in [0, n_reads)."""
    return sample_indices_many(seed, [int(site_id) & 0xFFFFFFFFFFFFFFFF], [n_reads], n_iters, n_samples)[0]


def sample_indices_mt19937(seed: int, n_reads: int, n_iters: int, n_samples: int = 20):
    """Indices the reference's ``np.random.choice(proba, n_iters*n_samples, replace=True)``
    uses when called first after ``np.random.seed(seed)`` (reference
    m6anet/utils/inference_utils.py:85, seeding at m6anet/scripts/inference.py:86).
    Legacy ``RandomState.choice`` with replacement draws ``randint(0, n, size)``.
    """
    rs = np.random.RandomState(seed)
    return rs.randint(0, n_reads, size=n_iters * n_samples).reshape(n_iters, n_samples).astype(np.int64)


# ---- bags WITHOUT replacement (the validate()-style literal MIL forward) --------------------------------
# The reference's evaluation datasets draw every bag with
# ``np.random.choice(len(features), self.min_reads, replace=False)`` (reference
# m6anet/utils/data_utils.py:213-214), once per site per validation pass
# (utils/training_utils.py:236-238).  Device specification ("Floyd bags"):
#
#   pass `it` of a site uses the same (site, block, lane) MWC64X stream and round as iteration `it` of the
#   inference stream above, one 32-bit word per draw (no paired regime):
#       for d in 0 .. k-1:   j = n_reads - k + d
#                            t = (word_d * (j + 1)) >> 32          uniform on [0, j]
#                            pick_d = j if t in {pick_0..pick_{d-1}} else t
#   (R. Floyd's sampling algorithm: the k picks are distinct and every k-subset is equally likely; the order inside
#    a bag carries no meaning for the pooling functions.)  Sites with n_reads < k have no bag (NaN on the device;
#    the reference's np.random.choice raises there, and its datasets drop such sites at :129).
def bag_indices_many(seed: int, site_ids, n_reads, n_iters: int, n_samples: int = 20) -> np.ndarray:
    """Device without-replacement bags for several sites -> int64 [n_sites, n_iters, n_samples]."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    site_ids = np.asarray(site_ids, dtype=np.uint64).reshape(-1)
    n_reads = np.asarray(n_reads, dtype=np.int64).reshape(-1)
    if np.any(n_reads < n_samples):
        raise ValueError("bag_indices: every site needs at least n_samples reads")
    S = len(site_ids)
    ipl, n_blocks = block_layout(n_iters)
    x, c = _streams(seed, site_ids, n_blocks)
    A = np.uint64(MWC_A)
    out = np.empty((S, n_blocks, ipl, 32, n_samples), dtype=np.int64)   # [S, block, round, lane, draw]
    nr = n_reads[:, None, None]
    for k in range(ipl):
        for d in range(n_samples):
            word = x ^ c
            t = A * x + c
            x, c = t & _MASK32, t >> _SHIFT32
            j = nr - n_samples + d
            cand = ((word * (j + 1).astype(np.uint64)) >> _SHIFT32).astype(np.int64)
            dup = (out[:, :, k, :, :d] == cand[..., None]).any(axis=-1) if d else np.zeros(cand.shape, dtype=bool)
            out[:, :, k, :, d] = np.where(dup, np.broadcast_to(j, cand.shape), cand)
    return out.reshape(S, n_blocks * ipl * 32, n_samples)[:, :n_iters, :]


def bag_indices(seed: int, site_id: int, n_reads: int, n_iters: int, n_samples: int = 20) -> np.ndarray:
    """Device without-replacement bags of one site -> int64 [n_iters, n_samples], distinct inside a bag."""
    return bag_indices_many(seed, [int(site_id) & 0xFFFFFFFFFFFFFFFF], [n_reads], n_iters, n_samples)[0]


def bag_indices_mt19937(rs: "np.random.RandomState", n_reads: int, n_samples: int = 20) -> np.ndarray:
    """The bag ``np.random.choice(n_reads, n_samples, replace=False)`` draws from the legacy stream `rs`
    (reference utils/data_utils.py:214): legacy ``RandomState.choice`` without replacement and without weights is
    ``permutation(n_reads)[:n_samples]``.  Consumes the stream exactly like the reference call."""
    return rs.permutation(int(n_reads))[:n_samples].astype(np.int64)
