// Host build of the SHIPPED device generator (m6anet_b200/csrc/m6a_rng.cuh: Philox4x32-10 seeding, MWC64X step, paired
// indices, Floyd bags) so that the CPU suite can hold the very source the kernel compiles to oracle/philox.py bit for
// bit.  Test infrastructure: built by tests/test_validate.py with g++, never part of the product library.
// The loops below mirror sample_indices_kernel (m6a_kernel.cu): one stream per (block, lane), rounds in order.
#include <stdint.h>

#include "../../m6anet_b200/csrc/m6a_rng.cuh"

using namespace m6a;

extern "C" void emul_sample(uint64_t seed, uint64_t site_id, uint32_t n_reads, int n_iters, int n_samples, int n_blocks,
                            int ipl, int without_replacement, int32_t* out) {
  for (int stream = 0; stream < n_blocks * 32; ++stream) {
    const int blk = stream >> 5, lane = stream & 31;
    Mwc64x g;
    g.seed(static_cast<uint32_t>(lane), static_cast<uint32_t>(blk), site_id, seed);
    for (int k = 0; k < ipl; ++k) {
      const long long it = (static_cast<long long>(blk) * ipl + k) * 32 + lane;
      if (it >= n_iters) break;
      if (without_replacement) {
        uint32_t pick[64];
        if (n_samples == 20) floyd_bag<20>(g, n_reads, n_samples, pick);   // both instantiations the kernel uses
        else floyd_bag<0>(g, n_reads, n_samples, pick);
        for (int s = 0; s < n_samples; ++s) out[it * n_samples + s] = static_cast<int32_t>(pick[s]);
        continue;
      }
      uint32_t pending = 0;
      for (int s = 0; s < n_samples; ++s) {
        uint32_t i;
        if (n_reads > kPairedMaxReads) i = mulhi_u32(g.next(), n_reads);
        else if ((s & 1) == 0) g.next_pair(n_reads, i, pending);
        else i = pending;
        out[it * n_samples + s] = static_cast<int32_t>(i);
      }
    }
  }
}
