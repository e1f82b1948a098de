// Device index stream: Philox4x32-10 seeds one MWC64X generator per (site, block, lane).
// Specification and NumPy twin: oracle/philox.py (compared bit for bit in
// tests/test_gpu_parity.py::test_device_index_stream_matches_oracle).
//
//   seeding   (w0, w1, _, _) = Philox4x32-10(ctr = (lane, block, site_lo, site_hi), key = (seed_lo, seed_hi))
//             x = w0;  c = (w1 * (A-1)) >> 32;  if (x == 0 && c == 0) x = 1
//   draw      word = x ^ c;  (c:x) = A * x + c                       A = 4294883355 (MWC64X, D. B. Thomas)
//   index     n_reads >  256: one index per word,  (word * n_reads) >> 32                     (bias <= n / 2^32)
//             n_reads <= 256: two indices per word, t = word * n_reads (64 bit):  i1 = t >> 32,
//                             i2 = ((t & 0xffffffff) * n_reads) >> 32                         (bias <= n^2 / 2^32 <= 1.6e-5)
//             an iteration consumes ceil(n_samples / 2) words in the paired regime (an odd last half is dropped)
//
// Why not Philox per draw: on sm_100a IMAD.WIDE/IMAD.HI issue at a quarter of the FFMA rate and
// occupy the FMA pipe (tools/microbench/rates.cu); Philox4x32-10 needs 20 of them per 4 words, the
// multiply-with-carry step needs one per word.
#pragma once
#include <stdint.h>

// The generator compiles for the host too (tests/native/bags_emul.cpp builds the same source with g++ and compares it
// with oracle/philox.py on the CPU); device code is unchanged by the host branches.
#if defined(__CUDACC__)
#define M6A_HD __host__ __device__ __forceinline__
#else
#define M6A_HD inline
#endif

namespace m6a {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kMwcA = 4294883355u;
constexpr int kMaxBlocks = 64;          // MC partial sums per site
constexpr uint32_t kPairedMaxReads = 256;   // sites with at most this many reads draw two indices per word
constexpr int kMinItersPerLane = 8;

struct Philox4 {
  uint32_t x, y, z, w;
};

M6A_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(kPhiloxM0) * c0;
    const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * c2;
    const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return Philox4{c0, c1, c2, c3};
}

M6A_HD uint32_t mulhi_u32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}

struct Mwc64x {
  uint32_t x, c;
  M6A_HD void seed(uint32_t lane, uint32_t block, uint64_t site, uint64_t key) {
    const Philox4 w = philox4x32_10(lane, block, static_cast<uint32_t>(site), static_cast<uint32_t>(site >> 32),
                                    static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32));
    x = w.x;
    c = mulhi_u32(w.y, kMwcA - 1u);
    if ((x | c) == 0u) x = 1u;
  }
  M6A_HD uint32_t next() {
    const uint32_t word = x ^ c;
#if defined(__CUDA_ARCH__)
    // (c:x) = A * x + c as one wide multiply (FMA pipe) plus an add-with-carry pair (ALU pipe).  The
    // fused form IMAD.WIDE Rd, x, A, (c,0) needs the addend in an aligned register pair, which costs
    // two extra moves per draw on the FMA pipe.
    uint32_t lo, hi;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(x), "r"(kMwcA));
    uint32_t nx, nc;
    asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=r"(nx), "=r"(nc) : "r"(lo), "r"(c), "r"(hi));
    x = nx;
    c = nc;
#else
    const uint64_t t = static_cast<uint64_t>(kMwcA) * x + c;
    x = static_cast<uint32_t>(t);
    c = static_cast<uint32_t>(t >> 32);
#endif
    return word;
  }
  // two indices from one word (paired regime): the wide product's high half is the first index, its low half is a
  // fresh uniform word for the second
  M6A_HD void next_pair(uint32_t n, uint32_t& i1, uint32_t& i2) {
#if defined(__CUDA_ARCH__)
    // the halves of the wide product are taken apart in PTX: written as (t >> 32) << 2 the address of q[i1] becomes
    // shift + mask + add (three instructions) instead of one LEA on the high register
    uint32_t lo;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(i1) : "r"(next()), "r"(n));
    i2 = mulhi_u32(lo, n);
#else
    const uint64_t t = static_cast<uint64_t>(next()) * n;
    i1 = static_cast<uint32_t>(t >> 32);
    i2 = mulhi_u32(static_cast<uint32_t>(t), n);
#endif
  }
};

// ---- bags without replacement (validate()-style literal MIL forward; specification: oracle/philox.py "Floyd bags") ----
// R. Floyd's algorithm: k distinct picks out of n, every k-subset equally likely, one 32-bit word per pick:
//   for d in 0..k-1:  j = n-k+d;  t = (word * (j+1)) >> 32;  pick_d = (t already picked) ? j : t
// Replaces np.random.choice(len(features), min_reads, replace=False) of the reference's evaluation datasets
// (utils/data_utils.py:213-214).  NS > 0: compile-time bag size (picks stay in registers); NS == 0: run-time k <= 64.
template <int NS, class Gen>
M6A_HD void floyd_bag(Gen& g, uint32_t n, int k_rt, uint32_t* pick) {
  const int k = NS > 0 ? NS : k_rt;
#pragma unroll
  for (int d = 0; d < k; ++d) {
    const uint32_t j = n - static_cast<uint32_t>(k) + static_cast<uint32_t>(d);
    const uint32_t t = mulhi_u32(g.next(), j + 1u);
    bool dup = false;
#pragma unroll
    for (int e = 0; e < d; ++e) dup |= (pick[e] == t);
    pick[d] = dup ? j : t;
  }
}

}  // namespace m6a
