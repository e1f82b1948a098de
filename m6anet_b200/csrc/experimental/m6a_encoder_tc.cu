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
//      A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo,   A_hi = the float32 value (the tensor core reads its top 19 bits),
//                                                 A_lo = A - trunc_tf32(A)  (exact in float32)
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

namespace m6a {
namespace tc {

struct alignas(128) Smem {
  float x[kK1 / 4][kTileM][4];       // A of Linear-1 (hi = the value itself)            8 KB
  float xlo[kK1 / 4][kTileM][4];     //                                                  8 KB
  float w1[kK1 / 4][kN1][4];         // B of Linear-1: [k-chunk][hidden unit][4]         10 KB
  float w1lo[kK1 / 4][kN1][4];
  float w2[kK2 / 4][kN2][4];         // B of Linear-2: [k-chunk][output][4]              20 KB
  float w2lo[kK2 / 4][kN2][4];
  float b2[kN2];
  float w3[kN2];
  float b3;
  uint32_t tmem_base;
  alignas(8) unsigned long long bar_l1;        // Linear-1 accumulators ready
  alignas(8) unsigned long long bar_stage[2];  // MMAs that read lo-staging buffer b have completed
  alignas(8) unsigned long long bar_l2;        // Linear-2 accumulators ready
};
static_assert(offsetof(Smem, xlo) % 128 == 0 && offsetof(Smem, w1) % 128 == 0 && offsetof(Smem, w2) % 128 == 0 &&
                  offsetof(Smem, w1lo) % 128 == 0 && offsetof(Smem, w2lo) % 128 == 0,
              "UMMA operands must start on a 128-byte core-matrix boundary");

size_t smem_bytes() { return sizeof(Smem); }

// ---- PTX wrappers -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (spins > (1u << 26)) __trap();   // a lost arrival must not hang the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_in_smem, uint32_t cols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_in_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t cols) {            // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// MMA completion -> one arrival on an mbarrier (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, cute/arch/mma_sm100_desc.hpp):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (between the two 16-byte k-chunks of one K-step),
//   [32,46) stride byte offset >> 4 (between 8-row core matrices), [46,48) version = 1, [61,64) layout = 0 (SWIZZLE_NONE)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}
// Instruction descriptor, kind::tf32, D = float32 (cute::UMMA::InstrDescriptor): [4,6) c_format = 1 (F32),
// [7,10) a_format = 2 (TF32), [10,13) b_format = 2, [15] a_major = 0 (K), [16] b_major = 0 (K), [17,23) N >> 3, [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane (base lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
      :
      : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]), "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float trunc_tf32(float f) { return __uint_as_float(__float_as_uint(f) & 0xFFFFE000u); }

// ---- the kernel ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
read_encoder_tc_kernel(const EncoderArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  // ---- one-time: weights -> shared memory (already in the UMMA byte layout), mbarriers, TMEM ----------------------------
  {
    const float4* src = reinterpret_cast<const float4*>(a.image);
    float4* dst = reinterpret_cast<float4*>(sm.w1);
    constexpr int n4 = (2 * kK1 * kN1 + 2 * kK2 * kN2) / 4;      // w1, w1lo, w2, w2lo are contiguous in both
    for (int i = tid; i < n4; i += kThreads) dst[i] = __ldg(src + i);
    if (tid < kN2) {
      sm.b2[tid] = a.image->b2[tid];
      sm.w3[tid] = a.image->w3[tid];
    }
    if (tid == 0) {
      sm.b3 = a.image->b3;
      mbar_init(&sm.bar_l1, 1);
      mbar_init(&sm.bar_stage[0], 1);
      mbar_init(&sm.bar_stage[1], 1);
      mbar_init(&sm.bar_l2, 1);
      fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&sm.tmem_base, kTmemCols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem = sm.tmem_base;                                   // lane 0, first allocated column
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;   // this warp's TMEM lane quadrant
  const uint32_t d1 = tmem + kColD1, lo_stage = tmem + kColLo, d2 = tmem + kColD2;
  constexpr uint32_t idesc1 = make_idesc(kTileM, kN1), idesc2 = make_idesc(kTileM, kN2);
  const uint32_t sx = smem_u32(sm.x), sxlo = smem_u32(sm.xlo), sw1 = smem_u32(sm.w1), sw1lo = smem_u32(sm.w1lo);
  const uint32_t sw2 = smem_u32(sm.w2), sw2lo = smem_u32(sm.w2lo);
  uint32_t ph_l1 = 0, ph_l2 = 0, ph_stage[2] = {0, 0};

  const long long n_tiles = (a.total_reads + kTileM - 1) / kTileM;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- stage: [x(9) | emb(6) | 1] of read r -> x, x_lo (UMMA K-major layout, one 16-byte k-chunk per store) ----------
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
#pragma unroll
    for (int j = 0; j < kK1 / 4; ++j) {
      float4 h = make_float4(in[4 * j], in[4 * j + 1], in[4 * j + 2], in[4 * j + 3]);
      float4 l = make_float4(h.x - trunc_tf32(h.x), h.y - trunc_tf32(h.y), h.z - trunc_tf32(h.z), h.w - trunc_tf32(h.w));
      *reinterpret_cast<float4*>(sm.x[j][tid]) = h;
      *reinterpret_cast<float4*>(sm.xlo[j][tid]) = l;
    }
    fence_proxy_async();
    tc_fence_before();       // the previous tile's tcgen05.ld of D1 / D2 are ordered before the MMAs that overwrite them
    __syncthreads();

    // ---- Linear-1: D1[128 x 160] = X . W1^T as 3 TF32 products per K-step -----------------------------------------------
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < kK1 / 8; ++ks) {
        const uint64_t ax = make_desc(sx + ks * kStepX, kLboX, kSbo);
        const uint64_t axlo = make_desc(sxlo + ks * kStepX, kLboX, kSbo);
        const uint64_t bw = make_desc(sw1 + ks * kStepW1, kLboW1, kSbo);
        const uint64_t bwlo = make_desc(sw1lo + ks * kStepW1, kLboW1, kSbo);
        mma_ss(d1, ax, bw, idesc1, ks > 0 ? 1u : 0u);
        mma_ss(d1, axlo, bw, idesc1, 1u);
        mma_ss(d1, ax, bwlo, idesc1, 1u);
      }
      tc_commit(&sm.bar_l1);
    }
    mbar_wait(&sm.bar_l1, ph_l1);
    ph_l1 ^= 1u;
    tc_fence_after();

    // ---- relu + split per chunk of 32 hidden units, Linear-2 on the chunk -----------------------------------------------
#pragma unroll 1
    for (int c = 0; c < kN1 / kChunk; ++c) {
      const int b = c & 1;
      uint32_t v[32], l[32];
      tmem_ld32(d1 + lane_base + c * kChunk, v);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float h = fmaxf(__uint_as_float(v[i]), 0.0f);
        v[i] = __float_as_uint(h);
        l[i] = __float_as_uint(h - trunc_tf32(h));
      }
      if (c >= 2) {            // the MMAs of chunk c-2 have finished reading staging buffer b
        mbar_wait(&sm.bar_stage[b], ph_stage[b]);
        ph_stage[b] ^= 1u;
        tc_fence_after();
      }
      tmem_st32(d1 + lane_base + c * kChunk, v);                    // A_hi of Linear-2, in place
      tmem_st32(lo_stage + lane_base + b * kChunk, l);              // A_lo
      tc_wait_st();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < kChunk / 8; ++ks) {
          const int kstep = c * (kChunk / 8) + ks;
          const uint64_t bw = make_desc(sw2 + kstep * kStepW2, kLboW2, kSbo);
          const uint64_t bwlo = make_desc(sw2lo + kstep * kStepW2, kLboW2, kSbo);
          const uint32_t a_hi = d1 + kstep * 8, a_lo = lo_stage + b * kChunk + ks * 8;
          mma_ts(d2, a_hi, bw, idesc2, kstep > 0 ? 1u : 0u);
          mma_ts(d2, a_lo, bw, idesc2, 1u);
          mma_ts(d2, a_hi, bwlo, idesc2, 1u);
        }
        tc_commit(&sm.bar_stage[b]);
        if (c == kN1 / kChunk - 1) tc_commit(&sm.bar_l2);
      }
    }
    // every commit is consumed exactly once, in order: chunks 3 (buffer 1) and 4 (buffer 0) are still outstanding
    mbar_wait(&sm.bar_stage[1], ph_stage[1]);
    ph_stage[1] ^= 1u;
    mbar_wait(&sm.bar_stage[0], ph_stage[0]);
    ph_stage[0] ^= 1u;
    mbar_wait(&sm.bar_l2, ph_l2);
    ph_l2 ^= 1u;
    tc_fence_after();

    // ---- epilogue: p = sigmoid(w3 . relu(D2 + b2) + b3) --------------------------------------------------------------------
    {
      uint32_t v[32];
      tmem_ld32(d2 + lane_base, v);
      tc_wait_ld();
      float z = sm.b3;
#pragma unroll
      for (int k = 0; k < kN2; ++k) z = fmaf(sm.w3[k], fmaxf(__uint_as_float(v[k]) + sm.b2[k], 0.0f), z);
      const float p = 1.0f / (1.0f + expf(-z));
      if (valid) a.read_prob[r] = p;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, kTmemCols);
}

cudaError_t launch_read_encoder_tc(const EncoderArgs& a, int n_sms, cudaStream_t stream) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  const int smem = static_cast<int>(sizeof(Smem));
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
      img->w1[k / 4][j][k % 4] = w;
      img->w1lo[k / 4][j][k % 4] = w - trunc_tf32_host(w);
    }
  }
  for (int n = 0; n < kN2; ++n)
    for (int k = 0; k < h1; ++k) {
      const float w = w2[static_cast<size_t>(n) * h1 + k];
      img->w2[k / 4][n][k % 4] = w;
      img->w2lo[k / 4][n][k % 4] = w - trunc_tf32_host(w);
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
struct m6a_tc_encoder {
  void* d_image;
  int n_kmer, emb_dim, n_sms;
};

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
