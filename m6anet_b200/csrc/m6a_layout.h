// Shared between the host-side packer (m6a_api.cu) and the kernel (m6a_kernel.cu).
#pragma once
#include <stdint.h>

namespace m6a {

constexpr int kNSig = 9;          // signal features per read (reference model_blocks/blocks.py:111: 3 * (2*1+1))
constexpr int kH2 = 32;           // second Linear width (reference m6anet.toml: output_channel = 32)
constexpr int kH1Max = 152;       // first Linear width limit (shipped: 150)
constexpr int kKmerPos = 3;       // five-mers per site (centre + 1 flank each side)
constexpr int kRowFloats = 44;    // per hidden unit j: w1[j,0..8], 3 pad, w2[0..31, j]
constexpr int kW2Off = 12;

// Image of the read-encoder weights as the kernel wants them in shared memory.  One
// cp.async.bulk moves it.  Per hidden unit j the 9 signal weights of Linear-1 (BatchNorm folded)
// sit next to column j of Linear-2, so the fused j-loop reads 11 consecutive float4 at a
// warp-uniform address (LDS.128 broadcast).
struct alignas(16) WeightImage {
  float l12[kH1Max][kRowFloats];  // 26,752 B
  float b2[kH2];
  float w3[kH2];
  float b3;
  int32_t h1;
  int32_t pad[2];
};
static_assert(sizeof(WeightImage) % 16 == 0, "bulk copy size must be a multiple of 16");

// Per-site constant of Linear-1: c[j] = b1[j] + sum_t W1[j, 9 + t*E .. ] . emb[kmer_t]  (the 3 k-mer ids
// are identical for every read of a site, reference utils/data_utils.py:223-224).  Stored in global
// memory (L2 resident) as three tables ctab[t][kmer][kH1Max]; table 0 carries b1.
//   c_site[j] = ctab[0][k0][j] + ctab[1][k1][j] + ctab[2][k2][j]
// For the signal-only topology n_kmer == 1, ctab[0][0] = b1 and tables 1,2 are zero.

struct DeviceModel {
  const WeightImage* image;   // device
  const float* ctab;          // device [3][n_kmer][kH1Max]
  int32_t n_kmer;
  int32_t h1;
  int32_t emb_dim;
};

}  // namespace m6a
