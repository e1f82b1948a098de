// Interface of the EXPERIMENTAL tensor-core read encoder (m6a_encoder_tc.cu); see the status note there.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../m6a_layout.h"

// handle of the experimental library's C entry points (m6a_tc_create / m6a_tc_read_probs_f32 / m6a_tc_mil_infer_f32)
struct m6a_tc_encoder {
  void* d_image;     // WeightImageTc on the device
  int n_kmer, emb_dim, n_sms;
};

namespace m6a {
namespace tc {

constexpr int kThreads = 128;      // 4 warps = the 4 TMEM lane quadrants
constexpr int kTileM = 128;        // reads per tile (TMEM lanes, UMMA M)
constexpr int kK1 = 16;            // Linear-1 inputs: 9 signal | 3 x emb_dim (<= 6) | bias column
constexpr int kN1 = 160;           // hidden units, padded from <= 152 to a multiple of the 32-column chunk
constexpr int kK2 = kN1;
constexpr int kN2 = kH2;           // 32
constexpr int kChunk = 32;         // hidden units per relu/split chunk = 4 K-steps of Linear-2
constexpr int kTmemCols = 256;     // D1 [0,160) | lo staging 2 x 32 [160,224) | D2 [224,256)
constexpr int kColD1 = 0, kColLo = kN1, kColD2 = kN1 + 2 * kChunk;
constexpr int kEmbMax = 4096 * 2;
static_assert(kColD2 + kN2 == kTmemCols, "TMEM column budget");

// UMMA K-major no-swizzle operand geometry ([k-chunk of 4 floats][row][4]): byte offsets the descriptors carry.
//   element (row, k) of an operand with R rows lives at (k / 4) * R * 16 + row * 16 + (k % 4) * 4
constexpr uint32_t kSbo = 128;                    // 8 rows x 16 bytes: next core matrix along M / N
constexpr uint32_t kLboX = kTileM * 16;           // next 16-byte k-chunk of the same rows
constexpr uint32_t kLboW1 = kN1 * 16;
constexpr uint32_t kLboW2 = kN2 * 16;
constexpr uint32_t kStepX = 2 * kLboX;            // one K-step (8 tf32 = 2 k-chunks)
constexpr uint32_t kStepW1 = 2 * kLboW1;
constexpr uint32_t kStepW2 = 2 * kLboW2;

// Instruction descriptor, kind::tf32, D = float32 (cute::UMMA::InstrDescriptor): [4,6) c_format = 1 (F32),
// [7,10) a_format = 2 (TF32), [10,13) b_format = 2, [15] a_major = 0 (K), [16] b_major = 0 (K), [17,23) N >> 3, [24,29) M >> 4
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// Weights as the kernel's shared memory holds them (K-major no-swizzle UMMA operands: [k-chunk of 4][row][4]);
// w* = trunc_tf32(w), w*lo = trunc_tf32(w - trunc_tf32(w)).  w1 rows = hidden units (BatchNorm folded), columns = [9 signal | 3 x emb | 0.. | b1 at 15].
struct alignas(16) WeightImageTc {
  float w1[kK1 / 4][kN1][4];
  float w1lo[kK1 / 4][kN1][4];
  float w2[kK2 / 4][kN2][4];
  float w2lo[kK2 / 4][kN2][4];
  float b2[kN2];
  float w3[kN2];
  float b3;
  float pad[3];
  float emb[kEmbMax];
};

struct EncoderArgs {
  const WeightImageTc* image;   // device
  const float* feats;           // [total_reads, 9]
  const int64_t* read_off;      // [n_sites + 1]
  const int32_t* kmer_idx;      // [n_sites, 3]
  float* read_prob;             // [total_reads]
  long long n_sites, total_reads;
  int n_kmer, emb_dim;
};

size_t smem_bytes();
bool pack_image(const float* emb, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                const float* b3, int n_kmer, int emb_dim, int h1, WeightImageTc* img);
cudaError_t launch_read_encoder_tc(const EncoderArgs& a, int n_sms, cudaStream_t stream);

}  // namespace tc
}  // namespace m6a
