// Shared between the host-side packer (m6a_api.cu) and the kernel (m6a_kernel.cu).
#pragma once
#include <stdint.h>

namespace m6a {

constexpr int kNSig = 9;          // signal features per read (reference model_blocks/blocks.py:111: 3 * (2*1+1))
constexpr int kH2 = 32;           // second Linear width (reference m6anet.toml: output_channel = 32)
constexpr int kH1Max = 152;       // first Linear width limit (shipped: 150)
constexpr int kMaxSamples = 64;   // reads per bag limit (the reference uses 20: utils/inference_utils.py:54, min_reads)
constexpr int kKmerPos = 3;       // five-mers per site (centre + 1 flank each side)
constexpr int kPairs = kH1Max / 2; // hidden units are processed two at a time (packed FFMA2)
constexpr int kPairFloats = 84;    // per pair p=(j0,j1): 9 x (w1[j0,k], w1[j1,k]), 2 pad, w2[0..31,j0], w2[0..31,j1]
constexpr int kW2Off0 = 20;        // float offset of w2[:, j0]
constexpr int kW2Off1 = 52;        // float offset of w2[:, j1]

// Image of the read-encoder weights as the kernel consumes them.  It is passed BY VALUE as a __grid_constant__
// kernel parameter (constant bank 0, 25.9 KB of the 32 KB limit): ptxas then streams it through the uniform
// datapath (LDCU.64 UR, c[0x0][UR+imm]) and FFMA2 takes the pair as a uniform-register operand, so the weights
// use neither shared-memory bandwidth nor vector registers (with -DM6A_WEIGHTS_CONST=0 the same image is staged in
// shared memory by one cp.async.bulk instead: 7 % slower, profiles/r01_tile_size_ab.txt).  Hidden units go in pairs (j0, j1) = (2p, 2p+1): the 9 signal weights of
// Linear-1 (BatchNorm folded) are interleaved as float2 (w1[j0,k], w1[j1,k]) so that one FFMA2 with
// the scalar x_k updates (h_j0, h_j1); columns j0 and j1 of Linear-2 follow as 2 x 16 float2 so that
// one FFMA2 with the scalar h_j updates two outputs.  The pair loop reads 21 consecutive float4 at a
// warp-uniform address (LDS.128 broadcast).  h1 odd => the last pair's j1 half is zero.
struct alignas(16) WeightImage {
  float pair[kPairs][kPairFloats];  // 25,536 B
  float b2[kH2];
  float w3[kH2];
  float b3;
  int32_t h1;
  int32_t n_pairs;
  int32_t pad;
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


// ---- tensor-core read encoder (m6a_kernel_tc.cu): tcgen05 kind::tf32, error-compensated 3xTF32 products ------------------
// Both Linear blocks run on the 5th-generation tensor cores, one tile of 128 reads (TMEM lanes) at a time:
//   Linear-1  D1[128 x 160] = [x(9) | emb(6) | 1] (K = 16) . W1a^T, W1a = [w1 | b1] (BatchNorm folded)
//   Linear-2  D2[128 x 32]  = relu(D1) (K = 160, A operand read from TMEM) . W2^T
// A 32-bit operand is truncated to TF32 by the hardware (measured, profiles/r02_tcgen05_probe.log), so every operand v is
// split into hi = RN_tf32(v) (stored with the low 13 mantissa bits zero) and lo = v - hi, and
//   A.B ~= A_hi.B_hi + (A_lo.B_hi + A_hi.B_lo)
// The tensor core adds into its float32 accumulator with truncation (a CPU model of exactly that reproduces the measured
// errors of round 1's experimental kernel to 10 %), so Linear-2 keeps the two small terms in their OWN accumulator and the
// epilogue adds main + correction once in round-to-nearest float32.
namespace tcx {
constexpr int kTileM = 128;        // reads per MMA tile (TMEM lanes, UMMA M)
constexpr int kK1 = 16;            // Linear-1 inputs: 9 signal | 3 x emb_dim (<= 6) | bias column
constexpr int kN1 = 160;           // hidden units, padded from <= 152 to 5 chunks of 32
constexpr int kN2 = kH2;           // 32
constexpr int kChunk = 32;         // hidden units per relu/split chunk = 4 K-steps of Linear-2
constexpr int kChunks = kN1 / kChunk;
constexpr int kEmbMax = 4096 * 2;  // embedding table floats kept in the image (n_kmer * emb_dim)

// Weights as the kernel's shared memory holds them: K-major no-swizzle UMMA operands [k-chunk of 4][row][4].
//   w1hi/w1lo  rows = hidden units, columns = [9 signal | 3 x emb | 0.. | b1 at 15]
//   w2s        rows 0..31 = W2_hi, rows 32..63 = W2_lo: one N = 64 MMA computes A_hi.W2_hi (main accumulator columns) and
//              A_hi.W2_lo (correction columns) at once; A_lo.W2_hi is an N = 32 MMA on rows 0..31 of the same operand
struct alignas(16) WeightImageTc {
  float w1hi[kK1 / 4][kN1][4];
  float w1lo[kK1 / 4][kN1][4];
  float w2s[kN1 / 4][2 * kN2][4];
  float b2[kN2];
  float w3[kN2];
  float b3;
  int32_t emb_dim;
  int32_t n_kmer;
  int32_t pad;
  float emb[kEmbMax];              // read from global memory (L1/L2 resident)
};
constexpr int kTcOperandBytes = (2 * kK1 * kN1 + kN1 * 2 * kN2) * 4;   // w1hi | w1lo | w2s, copied to shared memory
}  // namespace tcx

}  // namespace m6a
