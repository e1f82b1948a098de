// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11)
// and the product's index-stream specification.  oracle/philox.py is the NumPy twin; the two
// are compared bit for bit in tests/test_gpu_parity.py::test_device_philox_matches_oracle.
//
//   key   = (seed & 0xffffffff, seed >> 32)
//   ctr   = (call, iteration, site_id & 0xffffffff, site_id >> 32),  call = sample / 4
//   word  = philox4x32_10(ctr, key)[sample % 4]
//   index = (word * n_reads) >> 32
#pragma once
#include <stdint.h>

namespace m6a {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

struct Philox4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
    const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return Philox4{c0, c1, c2, c3};
}

}  // namespace m6a
