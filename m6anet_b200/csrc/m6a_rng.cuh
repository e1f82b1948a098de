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

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
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

struct Mwc64x {
  uint32_t x, c;
  __device__ __forceinline__ void seed(uint32_t lane, uint32_t block, uint64_t site, uint64_t key) {
    const Philox4 w = philox4x32_10(lane, block, static_cast<uint32_t>(site), static_cast<uint32_t>(site >> 32),
                                    static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32));
    x = w.x;
    c = __umulhi(w.y, kMwcA - 1u);
    if ((x | c) == 0u) x = 1u;
  }
  __device__ __forceinline__ uint32_t next() {
    const uint32_t word = x ^ c;
    // (c:x) = A * x + c as one wide multiply (FMA pipe) plus an add-with-carry pair (ALU pipe).  The
    // fused form IMAD.WIDE Rd, x, A, (c,0) needs the addend in an aligned register pair, which costs
    // two extra moves per draw on the FMA pipe.
    uint32_t lo, hi;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(x), "r"(kMwcA));
    uint32_t nx, nc;
    asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=r"(nx), "=r"(nc) : "r"(lo), "r"(c), "r"(hi));
    x = nx;
    c = nc;
    return word;
  }
  // two indices from one word (paired regime): the wide product's high half is the first index, its low half is a
  // fresh uniform word for the second
  __device__ __forceinline__ void next_pair(uint32_t n, uint32_t& i1, uint32_t& i2) {
    const uint64_t t = static_cast<uint64_t>(next()) * n;
    i1 = static_cast<uint32_t>(t >> 32);
    i2 = __umulhi(static_cast<uint32_t>(t), n);
  }
};

}  // namespace m6a
