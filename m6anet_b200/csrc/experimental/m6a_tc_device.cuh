// Device side of the EXPERIMENTAL tensor-core read encoder (tcgen05 + TMEM, 3xTF32), shared by the stand-alone encoder
// kernel (m6a_encoder_tc.cu) and the fused MIL kernel (m6a_fused_tc.cu).  Unvalidated on hardware -- see m6a_encoder_tc.cu.
//
// A CTA of W * 128 threads (W = 1 or 2 warps per TMEM lane quadrant) encodes one tile of 128 reads at a time:
//   tc_setup        weights -> shared memory (UMMA byte layout), mbarriers, 256 TMEM columns
//   tc_stage_row    thread t < 128 writes the 16 inputs of read t ([x(9) | emb(6) | 1]), truncated to TF32, and their residuals as the A
//                   operand of Linear-1
//   tc_encode_tile  Linear-1 (2 K-steps x 3 MMAs, SS) -> per chunk of 32 hidden units: relu, residual, back to TMEM,
//                   Linear-2 (4 K-steps x 3 MMAs, A from TMEM) -> p = sigmoid(w3 . relu(D2 + b2) + b3) for row t < 128
//   tc_teardown     TMEM release
// All CTA threads must call tc_setup / tc_encode_tile / tc_teardown together (they contain __syncthreads()).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "m6a_encoder_tc.h"

namespace m6a {
namespace tc {

struct alignas(128) TcSmem {
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
static_assert(offsetof(TcSmem, xlo) % 128 == 0 && offsetof(TcSmem, w1) % 128 == 0 && offsetof(TcSmem, w2) % 128 == 0 &&
                  offsetof(TcSmem, w1lo) % 128 == 0 && offsetof(TcSmem, w2lo) % 128 == 0,
              "UMMA operands must start on a 128-byte core-matrix boundary");

struct TcState {          // per-thread mbarrier phases and addresses
  uint32_t tmem, lane_base, ph_l1, ph_l2, ph_stage0, ph_stage1;
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void tc_mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(tc_smem_u32(bar)), "r"(parity)
        : "memory");
    if (spins > (1u << 26)) __trap();   // a lost arrival must not hang the GPU
  }
}
__device__ __forceinline__ void tc_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void tc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_in_smem, uint32_t cols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(dst_in_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t cols) {            // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// MMA completion -> one arrival on an mbarrier (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
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

// 32 lanes x N consecutive 32-bit columns: thread i of the warp gets lane (base lane + i)
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
      :
      : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(taddr)
      : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[N]) {
  static_assert(N == 16 || N == 32, "chunk share per warp");
  if constexpr (N == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
}
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&v)[N]) {
  if constexpr (N == 32) tmem_st32(taddr, v); else tmem_st16(taddr, v);
}

__device__ __forceinline__ float trunc_tf32(float f) { return __uint_as_float(__float_as_uint(f) & 0xFFFFE000u); }

// ---- CTA-level steps -----------------------------------------------------------------------------------------------------
// weights -> shared memory (already in the UMMA byte layout), mbarriers, TMEM.  All `n_threads` threads of the CTA.
__device__ __forceinline__ void tc_setup(TcSmem& sm, TcState& st, const WeightImageTc* image, int tid, int n_threads) {
  const float4* src = reinterpret_cast<const float4*>(image);
  float4* dst = reinterpret_cast<float4*>(sm.w1);
  constexpr int n4 = (2 * kK1 * kN1 + 2 * kK2 * kN2) / 4;      // w1, w1lo, w2, w2lo are contiguous in both
  for (int i = tid; i < n4; i += n_threads) dst[i] = __ldg(src + i);
  if (tid < kN2) {
    sm.b2[tid] = image->b2[tid];
    sm.w3[tid] = image->w3[tid];
  }
  if (tid == 0) {
    sm.b3 = image->b3;
    tc_mbar_init(&sm.bar_l1, 1);
    tc_mbar_init(&sm.bar_stage[0], 1);
    tc_mbar_init(&sm.bar_stage[1], 1);
    tc_mbar_init(&sm.bar_l2, 1);
    tc_fence_barrier_init();
  }
  if ((tid >> 5) == 0) tmem_alloc(&sm.tmem_base, kTmemCols);
  tc_fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  st.tmem = sm.tmem_base;                                               // lane 0, first allocated column
  st.lane_base = static_cast<uint32_t>(((tid >> 5) & 3) * 32) << 16;   // this warp's TMEM lane quadrant
  st.ph_l1 = st.ph_l2 = st.ph_stage0 = st.ph_stage1 = 0;
}

__device__ __forceinline__ void tc_teardown(TcSmem& sm, const TcState& st, int tid) {
  tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == 0) tmem_free(st.tmem, kTmemCols);
}

// thread `row` < 128: the 16 inputs of its read -> x, x_lo (UMMA K-major layout, one 16-byte k-chunk per store)
__device__ __forceinline__ void tc_stage_row(TcSmem& sm, int row, const float (&in)[kK1]) {
#pragma unroll
  for (int j = 0; j < kK1 / 4; ++j) {
    // hi is stored already truncated to TF32 (low 13 mantissa bits zero), so the split is exact whether the tensor core
    // truncates or rounds its 32-bit operands; lo = x - hi is exact in float32
    const float4 f = make_float4(in[4 * j], in[4 * j + 1], in[4 * j + 2], in[4 * j + 3]);
    const float4 h = make_float4(trunc_tf32(f.x), trunc_tf32(f.y), trunc_tf32(f.z), trunc_tf32(f.w));
    const float4 l = make_float4(f.x - h.x, f.y - h.y, f.z - h.z, f.w - h.w);
    *reinterpret_cast<float4*>(sm.x[j][row]) = h;
    *reinterpret_cast<float4*>(sm.xlo[j][row]) = l;
  }
}

// Encodes the 128 staged rows.  W warps share a TMEM lane quadrant and split every 32-column chunk (W = 1: 128 threads,
// W = 2: 256 threads).  Returns p of row (tid & 127) in the threads with tid < 128, 0 elsewhere.
// Every commit is consumed exactly once, in order, by every thread (strictly alternating mbarrier phases).
template <int W>
__device__ __forceinline__ float tc_encode_tile(TcSmem& sm, TcState& st, int tid) {
  static_assert(W == 1 || W == 2, "1 or 2 warps per lane quadrant");
  constexpr int kShare = kChunk / W;                      // columns of a chunk this warp handles
  constexpr uint32_t idesc1 = make_idesc(kTileM, kN1), idesc2 = make_idesc(kTileM, kN2);
  const int half = (tid >> 7) & (W - 1);                  // which share of the chunk (0 for W = 1)
  const uint32_t d1 = st.tmem + kColD1, lo_stage = st.tmem + kColLo, d2 = st.tmem + kColD2;
  const uint32_t sx = tc_smem_u32(sm.x), sxlo = tc_smem_u32(sm.xlo), sw1 = tc_smem_u32(sm.w1), sw1lo = tc_smem_u32(sm.w1lo);
  const uint32_t sw2 = tc_smem_u32(sm.w2), sw2lo = tc_smem_u32(sm.w2lo);

  tc_fence_proxy_async();  // the staged rows (generic-proxy stores) -> async proxy
  tc_fence_before();       // the previous tile's tcgen05.ld of D1 / D2 are ordered before the MMAs that overwrite them
  __syncthreads();

  // ---- Linear-1: D1[128 x 160] = X . W1^T as 3 TF32 products per K-step ---------------------------------------------------
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
  tc_mbar_wait(&sm.bar_l1, st.ph_l1);
  st.ph_l1 ^= 1u;
  tc_fence_after();

  // ---- relu + split per chunk of 32 hidden units, Linear-2 on the chunk ---------------------------------------------------
#pragma unroll 1
  for (int c = 0; c < kN1 / kChunk; ++c) {
    const int b = c & 1;
    uint32_t v[kShare], l[kShare];
    tmem_ld<kShare>(d1 + st.lane_base + c * kChunk + half * kShare, v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < kShare; ++i) {
      const float h = fmaxf(__uint_as_float(v[i]), 0.0f);
      const float hi = trunc_tf32(h);          // stored truncated: exact split under either operand-conversion behaviour
      v[i] = __float_as_uint(hi);
      l[i] = __float_as_uint(h - hi);
    }
    if (c >= 2) {            // the MMAs of chunk c-2 have finished reading staging buffer b
      tc_mbar_wait(&sm.bar_stage[b], b ? st.ph_stage1 : st.ph_stage0);
      if (b) st.ph_stage1 ^= 1u; else st.ph_stage0 ^= 1u;
      tc_fence_after();
    }
    tmem_st<kShare>(d1 + st.lane_base + c * kChunk + half * kShare, v);        // A_hi of Linear-2, in place
    tmem_st<kShare>(lo_stage + st.lane_base + b * kChunk + half * kShare, l);  // A_lo
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
  // chunks 3 (buffer 1) and 4 (buffer 0) are still outstanding
  tc_mbar_wait(&sm.bar_stage[1], st.ph_stage1);
  st.ph_stage1 ^= 1u;
  tc_mbar_wait(&sm.bar_stage[0], st.ph_stage0);
  st.ph_stage0 ^= 1u;
  tc_mbar_wait(&sm.bar_l2, st.ph_l2);
  st.ph_l2 ^= 1u;
  tc_fence_after();

  // ---- epilogue: p = sigmoid(w3 . relu(D2 + b2) + b3), one thread per row (the first warp of every quadrant) ------------
  float p = 0.0f;
  if (tid < kTileM) {
    uint32_t v[32];
    tmem_ld32(d2 + st.lane_base, v);
    tc_wait_ld();
    float z = sm.b3;
#pragma unroll
    for (int k = 0; k < kN2; ++k) z = fmaf(sm.w3[k], fmaxf(__uint_as_float(v[k]) + sm.b2[k], 0.0f), z);
    p = 1.0f / (1.0f + expf(-z));
  }
  return p;
}

}  // namespace tc
}  // namespace m6a
