// EXPERIMENTAL tensor-core read encoder for sm_100a (tcgen05 + TMEM, 3xTF32).  NOT on the product path.
//
// Status: written and cross-compiled in round 1 after the GPU budget was spent; it has never run on a GPU.  It is built
// into its OWN library (libm6a_encoder_tc.so, this directory's Makefile) that nothing in the product loads -- the product
// library stays free of tensor-core instructions as north_star asks (tests/test_cabi.py checks its SASS) -- and is
// reachable only through the three m6a_tc_* functions at the end of this file (tests/test_encoder_tc.py, GPU part skipped
// unless M6A_TEST_TC=1).  DESIGN.md section 8 explains why it exists: the CUDA-core encoder is bound by the FMA pipe
// (65 % busy, 13.8 of the 17.5 ms per pass) and north_star's premise "memory-bound, no tensor cores" does not hold for it.
//
// What it computes: read_prob[r] = sigmoid(w3 . relu(W2 relu(W1 [x | emb | 1]) + b2) + b3)   -- reference
// utils/inference_utils.py:35-37, model_blocks/blocks.py:126,204-205,65,249-255, pooling_blocks.py:52 -- with both Linear
// blocks on the tensor cores as error-compensated TF32 products (tools/tf32x3_feasibility.py: max |p - p_float64| = 5.9e-7
// with hardware truncation, 1xTF32 would be 9e-4):
//      A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo,   A_hi = trunc_tf32(A) (stored with the low 13 mantissa bits cleared, so the
//                                                 split is exact whether the hardware truncates or rounds its operands),
//                                                 A_lo = A - A_hi  (exact in float32)
//
// One CTA = 128 threads = one tile of 128 reads (TMEM lanes) at a time, persistent over tiles; 2 CTAs per SM (256 TMEM
// columns each):
//   stage   thread t gathers read t: 9 features + 6 embedding values + 1 (bias column) = K1 = 16, writes x and x_lo
//           into shared memory in the UMMA K-major no-swizzle layout (16-byte k-chunks: addr = chunk*R*16 + row*16)
//   L1      thread 0: 2 K-steps x 3 tcgen05.mma kind::tf32 (M128 N160 K8, A and B from shared memory) -> D1 = TMEM
//           columns [0,160), commit -> mbarrier
//   E1+L2   per chunk of 32 hidden units: tcgen05.ld (32 lanes x 32 columns per warp) -> relu, lo = h - trunc(h) ->
//           tcgen05.st h back IN PLACE (it is the A_hi operand of Linear-2, read from TMEM) and lo into a double-buffered
//           32-column staging area -> thread 0: 4 K-steps x 3 tcgen05.mma (M128 N32 K8, A from TMEM, B = W2 from shared
//           memory) -> D2 = TMEM columns [224,256), commit per staging buffer
//   E2      tcgen05.ld D2 -> + b2, relu, dot w3, sigmoid -> read_prob
// Every mbarrier wait is bounded (a lost arrival traps instead of hanging the GPU).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <new>

#include "../../../include/m6anet_b200.h"
#include "m6a_encoder_tc.h"
#include "m6a_tc_device.cuh"

namespace m6a {
namespace tc {

size_t smem_bytes() { return sizeof(TcSmem); }

// ---- the stand-alone encoder kernel: 128 threads, one tile of 128 consecutive reads at a time ---------------------------------
__global__ void __launch_bounds__(kThreads, 2)
read_encoder_tc_kernel(const EncoderArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x;
  TcState st;
  tc_setup(sm, st, a.image, tid, kThreads);

  const long long n_tiles = (a.total_reads + kTileM - 1) / kTileM;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- stage: [x(9) | emb(6) | 1] of read r ----------------------------------------------------------------------------
    const long long r = tile * kTileM + tid;
    const bool valid = r < a.total_reads;
    float in[kK1];
#pragma unroll
    for (int k = 0; k < kK1; ++k) in[k] = 0.0f;
    if (valid) {
#pragma unroll
      for (int k = 0; k < kNSig; ++k) in[k] = a.feats[r * kNSig + k];
      if (a.emb_dim > 0) {
        long long lo = 0, hi = a.n_sites;                 // site of read r: last s with read_off[s] <= r
        while (hi - lo > 1) {
          const long long mid = (lo + hi) >> 1;
          if (a.read_off[mid] <= r) lo = mid; else hi = mid;
        }
#pragma unroll
        for (int t = 0; t < kKmerPos; ++t) {
          int kid = a.kmer_idx[lo * kKmerPos + t];
          kid = min(max(kid, 0), a.n_kmer - 1);
          if (a.emb_dim == 2) {          // static register indices (the shipped models: Embedding(66, 2))
            in[kNSig + 2 * t] = __ldg(a.image->emb + kid * 2);
            in[kNSig + 2 * t + 1] = __ldg(a.image->emb + kid * 2 + 1);
          } else {                       // emb_dim == 1
            in[kNSig + t] = __ldg(a.image->emb + kid);
          }
        }
      }
      in[kK1 - 1] = 1.0f;                                 // bias column
    }
    tc_stage_row(sm, tid, in);
    const float p = tc_encode_tile<1>(sm, st, tid);
    if (valid) a.read_prob[r] = p;
  }
  tc_teardown(sm, st, tid);
}

cudaError_t launch_read_encoder_tc(const EncoderArgs& a, int n_sms, cudaStream_t stream) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  const int smem = static_cast<int>(sizeof(TcSmem));
  if (!attr_done[dev]) {
    e = cudaFuncSetAttribute(read_encoder_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done[dev] = true;
  }
  const long long n_tiles = (a.total_reads + kTileM - 1) / kTileM;
  if (n_tiles == 0) return cudaSuccess;
  long long grid = static_cast<long long>(n_sms) * 2;      // 2 CTAs per SM: 2 x 256 TMEM columns
  if (grid > n_tiles) grid = n_tiles;
  read_encoder_tc_kernel<<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

// ---- host: pack the weights into the shared-memory byte layout ------------------------------------------------------------
static float trunc_tf32_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u &= 0xFFFFE000u;
  memcpy(&f, &u, 4);
  return f;
}

// w1 [h1, 9 + 3E] (BatchNorm folded), b1 [h1], w2 [32, h1], b2, w3 [32], b3, emb [n_kmer, E]
bool pack_image(const float* emb, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                const float* b3, int n_kmer, int emb_dim, int h1, WeightImageTc* img) {
  if (h1 > kN1 || emb_dim < 0 || emb_dim > 2 || n_kmer * emb_dim > kEmbMax) return false;
  memset(img, 0, sizeof(*img));
  const int in1 = kNSig + kKmerPos * emb_dim;
  for (int j = 0; j < h1; ++j) {
    for (int k = 0; k < kK1; ++k) {
      float w = 0.0f;
      if (k < in1) w = w1[static_cast<size_t>(j) * in1 + k];
      else if (k == kK1 - 1) w = b1[j];
      const float hi = trunc_tf32_host(w);                      // operands are stored TF32-exact: the split does not
      img->w1[k / 4][j][k % 4] = hi;                            // depend on how the tensor core converts 32-bit inputs
      img->w1lo[k / 4][j][k % 4] = trunc_tf32_host(w - hi);
    }
  }
  for (int n = 0; n < kN2; ++n)
    for (int k = 0; k < h1; ++k) {
      const float w = w2[static_cast<size_t>(n) * h1 + k];
      const float hi = trunc_tf32_host(w);
      img->w2[k / 4][n][k % 4] = hi;
      img->w2lo[k / 4][n][k % 4] = trunc_tf32_host(w - hi);
    }
  for (int k = 0; k < kN2; ++k) {
    img->b2[k] = b2[k];
    img->w3[k] = w3[k];
  }
  img->b3 = b3[0];
  for (int i = 0; i < n_kmer * emb_dim; ++i) img->emb[i] = emb[i];
  return true;
}

}  // namespace tc
}  // namespace m6a

// ---- C entry points of the experimental library (plain C types; status codes of include/m6anet_b200.h) -----------------
// weights as for m6a_model_create (HOST pointers, BatchNorm folded into w1/b1); uploads the packed image to the current device
extern "C" int m6a_tc_create(const m6a_weights_t* w, m6a_tc_encoder** out) {
  if (!w || !out) return M6A_EINVAL;
  *out = nullptr;
  if (!w->w1 || !w->b1 || !w->w2 || !w->b2 || !w->w3 || !w->b3) return M6A_EINVAL;
  if (w->n_sig != m6a::kNSig || w->h2 != m6a::kH2 || w->h1 < 1) return M6A_EUNSUPPORTED;
  if (w->emb_dim > 0 && (!w->emb || w->n_kmer < 1)) return M6A_EINVAL;
  m6a::tc::WeightImageTc* img = new (std::nothrow) m6a::tc::WeightImageTc;
  if (!img) return M6A_ENOMEM;
  const int n_kmer = w->emb_dim > 0 ? w->n_kmer : 1;
  if (!m6a::tc::pack_image(w->emb, w->w1, w->b1, w->w2, w->b2, w->w3, w->b3, n_kmer, w->emb_dim, w->h1, img)) {
    delete img;
    return M6A_EUNSUPPORTED;
  }
  m6a_tc_encoder* h = new (std::nothrow) m6a_tc_encoder{nullptr, n_kmer, w->emb_dim, 0};
  int dev = 0;
  cudaError_t e = h ? cudaGetDevice(&dev) : cudaErrorMemoryAllocation;
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_image, sizeof(*img));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_image, img, sizeof(*img), cudaMemcpyHostToDevice);
  delete img;
  if (e != cudaSuccess) {
    if (h && h->d_image) cudaFree(h->d_image);
    delete h;
    return static_cast<int>(e);
  }
  *out = h;
  return M6A_OK;
}

extern "C" int m6a_tc_destroy(m6a_tc_encoder* h) {
  if (!h) return M6A_OK;
  cudaFree(h->d_image);
  delete h;
  return M6A_OK;
}

// DEVICE pointers; enqueues on `stream` of the current device; read_prob [total_reads]
extern "C" int m6a_tc_read_probs_f32(const m6a_tc_encoder* h, const float* feats, const int64_t* read_off,
                                     const int32_t* kmer_idx, int64_t n_sites, int64_t total_reads, float* read_prob,
                                     void* stream) {
  if (!h || n_sites < 0 || total_reads < 0) return M6A_EINVAL;
  if (total_reads == 0) return M6A_OK;
  if (!feats || !read_off || !read_prob || (h->emb_dim > 0 && !kmer_idx)) return M6A_EINVAL;
  m6a::tc::EncoderArgs a;
  a.image = static_cast<const m6a::tc::WeightImageTc*>(h->d_image);
  a.feats = feats;
  a.read_off = read_off;
  a.kmer_idx = kmer_idx;
  a.read_prob = read_prob;
  a.n_sites = n_sites;
  a.total_reads = total_reads;
  a.n_kmer = h->n_kmer;
  a.emb_dim = h->emb_dim;
  const cudaError_t e = m6a::tc::launch_read_encoder_tc(a, h->n_sms, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? M6A_OK : static_cast<int>(e);
}

// ---- CPU-checkable hooks (no CUDA call): the packed image and the operand geometry the descriptors are built from ----
extern "C" int64_t m6a_tc_image_bytes(void) { return static_cast<int64_t>(sizeof(m6a::tc::WeightImageTc)); }

extern "C" int m6a_tc_debug_image(const m6a_weights_t* w, void* out, int64_t out_bytes) {
  if (!w || !out || out_bytes < static_cast<int64_t>(sizeof(m6a::tc::WeightImageTc))) return M6A_EINVAL;
  const int n_kmer = w->emb_dim > 0 ? w->n_kmer : 1;
  return m6a::tc::pack_image(w->emb, w->w1, w->b1, w->w2, w->b2, w->w3, w->b3, n_kmer, w->emb_dim, w->h1,
                             static_cast<m6a::tc::WeightImageTc*>(out))
             ? M6A_OK
             : M6A_EUNSUPPORTED;
}

// out[16]: byte offsets of w1, w1lo, w2, w2lo, b2, w3, b3, emb in the image; SBO; LBO and K-step bytes of W1 and W2; the two
// instruction descriptors (M128 N160 / M128 N32)
extern "C" int m6a_tc_geometry(int64_t* out) {
  using m6a::tc::WeightImageTc;
  if (!out) return M6A_EINVAL;
  out[0] = offsetof(WeightImageTc, w1);
  out[1] = offsetof(WeightImageTc, w1lo);
  out[2] = offsetof(WeightImageTc, w2);
  out[3] = offsetof(WeightImageTc, w2lo);
  out[4] = offsetof(WeightImageTc, b2);
  out[5] = offsetof(WeightImageTc, w3);
  out[6] = offsetof(WeightImageTc, b3);
  out[7] = offsetof(WeightImageTc, emb);
  out[8] = m6a::tc::kSbo;
  out[9] = m6a::tc::kLboW1;
  out[10] = m6a::tc::kStepW1;
  out[11] = m6a::tc::kLboW2;
  out[12] = m6a::tc::kStepW2;
  out[13] = m6a::tc::make_idesc(m6a::tc::kTileM, m6a::tc::kN1);
  out[14] = m6a::tc::make_idesc(m6a::tc::kTileM, m6a::tc::kN2);
  out[15] = m6a::tc::kTmemCols;
  return M6A_OK;
}
