// tcgen05 / TMEM / mbarrier building blocks of the tensor-core read encoder (m6a_kernel_tc.cu), sm_100a inline PTX.
// Verified on a B200 by tools/microbench/tcgen05_probe.cu (profiles/r02_tcgen05_probe.log): TMEM st/ld round trip,
// kind::tf32 SS and TS products, K-step descriptor advance, and that 32-bit operands are TRUNCATED to TF32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace m6a {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a lost arrival must not hang the GPU box.  `site` identifies the wait in the trap record.
__device__ int* g_trap_record = nullptr;   // optional mapped host buffer (m6a_tc_set_trap_record): survives the trap
#ifndef M6A_WAIT_HINT_NS
#define M6A_WAIT_HINT_NS 0        // > 0: mbarrier.try_wait with this suspend-time hint (ns)
#endif
#ifndef M6A_WAIT_SLEEP_NS
#define M6A_WAIT_SLEEP_NS 0       // > 0: __nanosleep between groups of polls
#endif
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar_addr, uint32_t parity) {
  uint32_t done;
#if M6A_WAIT_HINT_NS > 0
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar_addr), "r"(parity), "r"(static_cast<uint32_t>(M6A_WAIT_HINT_NS))
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar_addr), "r"(parity)
      : "memory");
#endif
  return done;
}
// Bounded wait.  ncu of the first version (one poll, one counter update and one limit test per trip) showed a third of all
// issued instructions in this loop, so the polls come four to a trip and the bookkeeping once per trip.
__device__ __noinline__ void mbar_wait_timeout(int site, uint32_t parity) {
  if (g_trap_record != nullptr && site >= 0 && site < 16) {
    int* rec = g_trap_record + 4 * site;
    rec[0] = site;
    rec[1] = static_cast<int>(blockIdx.x);
    rec[2] = static_cast<int>(threadIdx.x);
    rec[3] = static_cast<int>(parity);
    __threadfence_system();
  }
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity, int site) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try_wait(addr, parity)) return;
  for (uint32_t trips = 0;; ++trips) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (mbar_try_wait(addr, parity)) return;
#if M6A_WAIT_SLEEP_NS > 0
    __nanosleep(M6A_WAIT_SLEEP_NS);
#endif
    if (trips >= (1u << 16)) {                       // ~1 s (a trip is four try_waits of up to ~4 us): every stuck wait leaves its
      if (trips == (1u << 16)) mbar_wait_timeout(site, parity);   // record, the first one traps a few seconds later
      if (trips > (1u << 18)) __trap();
    }
  }
}

// ---- proxies and tcgen05 fences -----------------------------------------------------------------------------------------
// generic-proxy writes to shared memory -> visible to the async proxy (the tensor core reads its operands through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_in_smem, uint32_t cols) {   // one full warp; power of two >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_in_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t cols) {            // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// completion of all MMAs issued so far by this thread -> one arrival on an mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): [0,14) start >> 4, [16,30) leading
// byte offset >> 4 (between the two 16-byte k-chunks of one K-step), [32,46) stride byte offset >> 4 (between 8-row core
// matrices), [46,48) version = 1, [61,64) layout = 0 (SWIZZLE_NONE).  Element (row, k) of an operand with R rows lives at
// (k / 4) * R * 16 + row * 16 + (k % 4) * 4.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}
// Instruction descriptor, kind::tf32, D = float32 (cute::UMMA::InstrDescriptor): [4,6) c_format = 1 (F32), [7,10) a_format
// = 2 (TF32), [10,13) b_format = 2, [15] a_major = 0 (K), [16] b_major = 0 (K), [17,23) N >> 3, [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// One lane of a converged warp.  tcgen05.mma / tcgen05.commit take their operands from the uniform datapath: issued under this
// predicate they compile to `ELECT; @P UTCHMMA`, whereas under `if (lane == 0)` ptxas wraps EVERY instruction in a
// loop over the active lanes (ELECT / BRA.U.ANY, ~9 dependent instructions and a branch per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[smem] . B[smem]^T
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane (quadrant base + i).  A warp may only touch the
// TMEM lane quadrant 32 * (warp_id % 4).
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

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

// 4-byte asynchronous global -> shared copy (LDGSTS): feature rows are only 4-byte aligned (36-byte rows)
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src_global) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_global) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// MMA forms with the accumulate flag fixed at compile time (no predicate set-up in the issue loop)
__device__ __forceinline__ void mma_ss_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  asm volatile("tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, 1;" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ts_acc(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
  asm volatile("tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, 1;" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc) : "memory");
}

// named barrier among `count` threads (ids 1.. ; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

}  // namespace tc
}  // namespace m6a
