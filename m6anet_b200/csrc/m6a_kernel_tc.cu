// Fused MIL-inference kernel for sm_100a with the read encoder on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same contract and the same results (index streams, summation order) as mil_infer_kernel (m6a_kernel.cu); what changes is
// WHERE the work runs.  ncu of the CUDA-core kernel: DRAM 1.4 % busy, FMA pipe 65 % -- the path is bound by FP32 issue, and
// 11.4 of its 17.5 ms per pass are the two Linear blocks (profiles/r02_tiles_static_1m_it1.json).  Here they are two
// tcgen05.mma chains per tile of 128 reads (error-compensated 3xTF32, m6a_layout.h), and the CUDA cores keep only the
// relu / operand split, the sigmoid and the Monte-Carlo pooling.
//
// One persistent CTA per SM, 1024 threads, warp-specialised; the roles run concurrently on DIFFERENT tiles / slabs
// (slab = the <= 64 consecutive sites of one slice of a read-balanced tile, as in m6a_kernel.cu) and meet only at mbarriers:
//   warp 31         MMA issuer.  The whole warp runs the loop, one elect.sync lane issues: Linear-1 of tile t (6 MMAs M128 N160
//                   K8, A and B from shared memory), then per 32-column chunk of tile t-1: 8 MMAs (A from TMEM: N64
//                   main|correction + N32 correction) and the tcgen05.commit that hands the chunk's A_lo slot back
//   warps 20..27    E1, two per TMEM lane quadrant (each takes 16 of a chunk's 32 columns), thread = read: per chunk
//                   tcgen05.ld D1 -> relu -> hi/lo split -> tcgen05.st hi IN PLACE (A operand of Linear-2) and lo into a
//                   2-slot ring -> mbarrier -> MMA issuer
//   warps 12..19    staging + E2, two groups of 4 warps taking the tiles alternately, thread = read:
//                     X   cp.async prefetch of [x(9) | emb(6) | 1] -> RN_tf32 split -> shared memory (UMMA K-major layout)
//                     E2  tcgen05.ld D2 (main + correction of two accumulator groups) -> +b2, relu, . w3, sigmoid ->
//                         read_prob (HBM), q = 1 - p (shared memory of the slab slot), threshold count; last tile of a
//                         slab -> slab_full
//   warps 0..11, 28..30  Monte-Carlo pooling of a FINISHED slab (m6a_mc.cuh: a warp takes whole sites, the blocks of 256
//                   iterations of a site are its interleaved chains; Philox-seeded MWC64X lane streams), then site_prob /
//                   mod_count and slab_empty
// TMEM (512 columns): D1[2] 2 x 160 | A_lo ring 2 x 32 | D2 = 2 accumulator groups x (32 main + 32 correction).
// Every mbarrier wait is bounded (a lost arrival leaves a record in mapped host memory and traps instead of hanging the GPU).
// Design notes, measurements and the experiments behind the constants below: DESIGN.md sections 4a and 7.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "m6a_kernel.h"
#include "m6a_mc.cuh"
#include "m6a_rng.cuh"
#include "m6a_tc.cuh"

namespace m6a {
namespace tcx {

using namespace m6a::tc;

constexpr int kTcThreads = 1024;
#ifndef M6A_TC_LAYOUT
#define M6A_TC_LAYOUT 1     // 1: the latency-critical roles sit on the highest warp ids (12.73 vs 12.84 ms on the headline job)
#endif
#if M6A_TC_LAYOUT == 0
constexpr int kMmaWarp = 1;
constexpr int kE1Warp0 = 4;                  // warps 4..11 : E1 (relu / split), two per TMEM lane quadrant (quadrant = warp % 4)
constexpr int kSeWarp0 = 12;                 // warps 12..19: staging of X + E2 (sigmoid, outputs), two per quadrant
#else                                        // the latency-critical roles on the highest warp ids (scheduler priority experiment)
constexpr int kMmaWarp = 31;
constexpr int kE1Warp0 = 20;
constexpr int kSeWarp0 = 12;
#endif
constexpr int kRoleWarps = 8;
constexpr int kRoleThreads = kRoleWarps * 32;   // 256
constexpr int kMcWarps = 15;                 // warps 0..11, 28..30 (layout 1)
// 1: Monte-Carlo pooling with lanes = sites (site-interleaved q table, conflict-free loads) where the slab fits.  Measured on
// B200 (1 M x 50 x 1000): 11.15 ms with it, 10.51 ms without -- the bank conflicts of the row-order table are not what
// limits the pooling; kept as an experiment (parity-checked, bit-identical site sums).
#ifndef M6A_TC_LANES_SITES
#define M6A_TC_LANES_SITES 0
#endif
#ifndef M6A_MC_CHAINS
#define M6A_MC_CHAINS 2     // measured on B200 (1 M x 50 x 1000, final pipeline): 1 / 2 / 3 / 4 chains 10.66 / 10.47 / 11.73 / 11.48 ms
#endif
constexpr int kMcChains = M6A_MC_CHAINS;     // blocks of a site a Monte-Carlo warp interleaves
#ifndef M6A_TC_SLOTS
#define M6A_TC_SLOTS 3
#endif
constexpr int kSlots = M6A_TC_SLOTS;                    // slab slots in shared memory (q table, offsets, counters)
constexpr int kSlabSites = kSitesPerTileMax; // 64
constexpr int kTmemCols = 512;
constexpr uint32_t kColD1 = 0, kColALo = 2 * kN1, kColD2 = 2 * kN1 + 2 * kChunk;   // 0 | 320 | 384
constexpr int kGroups = 2;                   // Linear-2 accumulator groups (chunks 0..2 | 3..4), each [main 32 | corr 32]
constexpr int kGroup1Chunk = 3;              // first chunk of the second group
#ifndef M6A_D2_SPLIT
#define M6A_D2_SPLIT 0      // 1: D2 handed to E2 / back to the MMA issuer per accumulator group; 0: as a whole
#endif
static_assert(kColD2 + kGroups * 2 * kN2 == kTmemCols, "TMEM column budget");
constexpr int kBarSe = 1, kBarE1 = 3;        // named barriers: staging / E2 groups (1, 2), E1 role (3)
constexpr int kHalfCols = kChunk / 2;        // columns of a chunk per warp
constexpr int kCtrlRing = 8;                 // per-tile control words (the staging warps run at most 4 tiles ahead of E1)

// UMMA operand geometry (byte offsets the descriptors carry; see make_desc)
constexpr uint32_t kSbo = 128;
constexpr uint32_t kLboX = kTileM * 16, kStepX = 2 * kLboX;
constexpr uint32_t kLboW1 = kN1 * 16, kStepW1 = 2 * kLboW1;
constexpr uint32_t kLboW2 = 2 * kN2 * 16, kStepW2 = 2 * kLboW2;

constexpr int kQCapT = kTcQCap;              // q entries of a slab slot (24 KB)
constexpr int kLaneBlocksMax = 4;            // lanes = sites pooling: blocks of iterations per site it covers (n_iters <= 1024)
constexpr int kOctets = 8;                   // a block's 32 lane streams in 8 groups of 4 = the b_k level of the butterfly sum

struct SlabMeta {
  long long s0;      // first site of the slab (shard-local)
  long long r0;      // first feature row
  int ns, nr;        // sites, rows
  int stop;          // 1: no more slabs for this CTA
  int max_n;         // most reads of a site of the slab
  int mode;          // 1: q stored site-interleaved (lanes = sites pooling), 0: q in row order, -1: q not in shared memory
  int pad;
};

struct alignas(128) TcSmem {
  float w1hi[kK1 / 4][kN1][4];                  // 10 KB   B of Linear-1
  float w1lo[kK1 / 4][kN1][4];                  // 10 KB
  float w2s[kN1 / 4][2 * kN2][4];               // 40 KB   B of Linear-2 (rows 0..31 hi, 32..63 lo)
  float x[2][2][kK1 / 4][kTileM][4];            // 32 KB   A of Linear-1: [buffer][hi, lo][k-chunk][row][4]
  float fbuf[2][kK1][kTileM];                   // 16 KB   prefetched inputs of a group's next tile (cp.async), input-major:
                                                //         [x(9) | emb | 1][row] -- lanes = rows, no bank conflicts
#if M6A_D2_SPLIT
  float g0[2][kN2][kTileM];                     // 32 KB   E2: (main + corr) of accumulator group 0, per E2 group, column-major
                                                //         (thread = row: private, conflict-free) until group 1 is complete
#endif
  float q[kSlots][kQCapT];                      // 72 KB   q = 1 - p of a slab (row order, or [group][entry][32 sites])
#if M6A_TC_LANES_SITES
  float part[kSlots][2][kLaneBlocksMax][kOctets][32];   // 24 KB  lanes = sites pooling: b_k partial sums per (group, block)
#endif
  float b2[kN2];
  float w3[kN2];
  float b3;
  uint32_t tmem_base;
  int x_ctrl[kCtrlRing];                        // per tile: 1 = staged, 0 = stop (written before the x_full arrivals)
  long long next_tile;                          // broadcast of the dynamic tile counter
  SlabMeta meta[kSlots];
  int roff[kSlots][kSlabSites + 1];
  int cnt[kSlots][kSlabSites];
  int kid[kSlots][kSlabSites][kKmerPos];
  alignas(8) unsigned long long x_full[2];      // staging (256) -> MMA, E1 : operand X[b] staged
  alignas(8) unsigned long long l1_done[2];     // MMA commit    -> E1, staging : D1[b] ready / X[b] consumed
  alignas(8) unsigned long long a_full[2];      // E1 (256)      -> MMA : chunk staged in D1 (hi) / A_lo slot
  alignas(8) unsigned long long a_free[2];      // MMA commit    -> E1 : A_lo slot consumed
  // D2 is handed over per accumulator group: group 0 (chunks 0..2) is complete -- and is read out by E2 -- while the MMAs
  // of chunks 3..4 still run, so Linear-2 of the next tile never waits for the read-out of a whole D2.
  alignas(8) unsigned long long d2_full[kGroups][2];   // MMA commit -> E2 group (tile parity) : accumulator group complete.
                                                // One barrier per E2 group: a parity wait must never be posted a phase
                                                // early, and the groups take turns
  alignas(8) unsigned long long d2_free[kGroups];      // E2 (128)   -> MMA : accumulator group read out
  int done[kSlots];                             // MC warps finished with the slab of a slot
  int maxn_scratch[2][4];
  alignas(8) unsigned long long slab_full[kSlots];   // E2 (128) -> MC
  alignas(8) unsigned long long slab_empty[kSlots];  // MC (1)   -> staging
  alignas(8) unsigned long long hdr_ready[kSlots];   // staging group that opened the slab (128) -> the other group
};
static_assert(offsetof(TcSmem, w1lo) % 128 == 0 && offsetof(TcSmem, w2s) % 128 == 0 && offsetof(TcSmem, x) % 128 == 0,
              "UMMA operands must start on a 128-byte core-matrix boundary");
static_assert(offsetof(TcSmem, w1hi) == 0 && offsetof(TcSmem, w2s) + sizeof(float) * kN1 / 4 * 2 * kN2 * 4 == kTcOperandBytes,
              "operand block is one contiguous copy of the image head");

size_t tc_smem_bytes(int /*n_blocks*/) { return sizeof(TcSmem); }

// ---- optional phase profile (-DM6A_TC_PROFILE=1): cycles per phase of one thread per role of block 0 ------------------------
#ifndef M6A_TC_PROFILE
#define M6A_TC_PROFILE 0
#endif
#if M6A_TC_PROFILE
__device__ unsigned long long g_prof[40];
#define PROF_DECL unsigned long long prof_t0 = clock64(); unsigned long long prof_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define PROF(i) do { const unsigned long long prof_t1 = clock64(); prof_acc[i] += prof_t1 - prof_t0; prof_t0 = prof_t1; } while (0)
#define PROF_STORE(base, cond) do { if ((cond) && blockIdx.x == 0) for (int pi = 0; pi < 10; ++pi) g_prof[(base) + pi] = prof_acc[pi]; } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_STORE(base, cond)
#endif

// ---- ablation switches for timing experiments only (results are WRONG when any is set) ---------------------------------------
#ifndef M6A_ABL
#define M6A_ABL 0      // bit 0: no proxy fence; 1: E1 without TMEM traffic / math; 2: no MMAs issued; 3: E2 without TMEM load / math;
#endif                 // bit 4: staging without cp.async / split; 5: MC items skipped

// wait sites (trap record)
enum WaitSite : int {
  kWaitXFull = 1, kWaitL1Done, kWaitAFull, kWaitAFree, kWaitD2Full, kWaitD2Free, kWaitSlabFull, kWaitSlabEmpty,
  kWaitXFullE1, kWaitL1DoneSe, kWaitHdr
};

struct TileInfo {      // one MMA tile of 128 rows, as seen by the staging / E2 thread that owns row `row`
  long long grow;      // global feature row of this thread
  int slot;            // slab slot
  int lr;              // slab-local row
  int site_l;          // slab-local site of the row
  bool exists, valid, last;
};

__device__ __forceinline__ float rn_tf32(float v) {     // round to nearest (ties away), low 13 mantissa bits zero
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}

// Role-wide wait on an mbarrier: ONE lane of the role's first warp polls it, the other warps of the role sleep in a
// hardware barrier.  (Eight warps polling try_wait cost ~30 % of the SM's issue slots while the pipeline waits for the
// Monte-Carlo warps -- ncu, profiles/r02_tc_abl30_*: the Monte-Carlo warps lose exactly those slots.)
#ifndef M6A_ROLE_WAIT
#define M6A_ROLE_WAIT 0     // measured: the extra barrier hop costs more than the polls it saves (13.65 vs 12.67 ms)
#endif
__device__ __forceinline__ void role_wait(unsigned long long* bar, uint32_t parity, int site, bool leader_warp, int lane,
                                          int bar_id, int n_threads) {
#if M6A_ROLE_WAIT
  if (leader_warp) {
    if (lane == 0) mbar_wait(bar, parity, site);
    __syncwarp();
  }
  named_bar_sync(bar_id, n_threads);
#else
  mbar_wait(bar, parity, site);
#endif
}

template <int NS>
__global__ void __launch_bounds__(kTcThreads, 1)
mil_infer_tc_kernel(const KernelArgs a, const WeightImageTc* __restrict__ image) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_blocks = a.n_blocks, ipl = a.iters_per_lane;

  // ---- one-time setup: operands -> shared memory, barriers, TMEM -----------------------------------------------------------
  {
    const float4* src = reinterpret_cast<const float4*>(image);
    float4* dst = reinterpret_cast<float4*>(&sm.w1hi[0][0][0]);
    for (int i = tid; i < kTcOperandBytes / 16; i += kTcThreads) dst[i] = __ldg(src + i);
    if (tid < kN2) {
      sm.b2[tid] = image->b2[tid];
      sm.w3[tid] = image->w3[tid];
    }
    if (tid == 0) {
      sm.b3 = image->b3;
      for (int i = 0; i < 2; ++i) {
        mbar_init(&sm.x_full[i], kTileM);
        mbar_init(&sm.l1_done[i], 1);
        mbar_init(&sm.a_full[i], kRoleThreads);
        mbar_init(&sm.a_free[i], 1);
      }
      for (int i = 0; i < kGroups; ++i) {
        mbar_init(&sm.d2_full[i][0], 1);
        mbar_init(&sm.d2_full[i][1], 1);
        mbar_init(&sm.d2_free[i], kTileM);
      }
      for (int i = 0; i < kSlots; ++i) {
        mbar_init(&sm.slab_full[i], kTileM);
        mbar_init(&sm.slab_empty[i], 1);
        mbar_init(&sm.hdr_ready[i], kTileM);
        sm.done[i] = 0;
      }
      fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&sm.tmem_base, kTmemCols);
    fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
  }
  const uint32_t tmem = sm.tmem_base;

  if (warp == kMmaWarp) {
    // ======================================== MMA issuer ===================================================================
    // The whole warp runs the loop (all state is warp-uniform) and one elected lane issues: see elect_one().
    {
      constexpr uint32_t idesc1 = make_idesc(kTileM, kN1), idesc64 = make_idesc(kTileM, 2 * kN2), idesc32 = make_idesc(kTileM, kN2);
      // descriptors differ only in the start-address field (bits [0,14) = address >> 4): build once, add offsets
      const uint64_t dx0 = make_desc(smem_u32(sm.x[0][0]), kLboX, kSbo);
      const uint64_t dw1hi = make_desc(smem_u32(sm.w1hi), kLboW1, kSbo), dw1lo = make_desc(smem_u32(sm.w1lo), kLboW1, kSbo);
      const uint64_t dw2 = make_desc(smem_u32(sm.w2s), kLboW2, kSbo);
      constexpr uint64_t kXBuf = sizeof(sm.x[0]) >> 4, kXLo = sizeof(sm.x[0][0]) >> 4;
      uint32_t g = 0;                     // chunk counter (A_lo slot = g & 1)
      bool stop = false;
      PROF_DECL;
      for (uint32_t t = 0;; ++t) {
        const uint32_t b = t & 1u;
        // ---- Linear-1 of tile t ------------------------------------------------------------------------------------------
        if (!stop) {
          mbar_wait(&sm.x_full[b], (t >> 1) & 1u, kWaitXFull);
          stop = sm.x_ctrl[t % kCtrlRing] == 0;
        }
        PROF(0);
        if (!stop) {
          fence_after();
          const uint64_t ax = dx0 + b * kXBuf, axlo = ax + kXLo;
          const uint32_t d1 = tmem + kColD1 + b * kN1;
          if (elect_one()) {       // the commit is issued by the lane that issued the MMAs it tracks
            if (!(M6A_ABL & 4)) {
              mma_ss(d1, ax, dw1hi, idesc1, 0u);
              mma_ss_acc(d1, axlo, dw1hi, idesc1);
              mma_ss_acc(d1, ax, dw1lo, idesc1);
              mma_ss_acc(d1, ax + (kStepX >> 4), dw1hi + (kStepW1 >> 4), idesc1);
              mma_ss_acc(d1, axlo + (kStepX >> 4), dw1hi + (kStepW1 >> 4), idesc1);
              mma_ss_acc(d1, ax + (kStepX >> 4), dw1lo + (kStepW1 >> 4), idesc1);
            }
            mma_commit(&sm.l1_done[b]);
          }
        } else {
          if (elect_one()) mbar_arrive(&sm.l1_done[b]);           // wakes E1, which then reads the stop word
        }
        __syncwarp();
        PROF(1);
        // ---- Linear-2 of tile t-1, chunk by chunk as E1 stages them -----------------------------------------------------------
        if (t > 0) {
          const uint32_t u = t - 1, j = u & 1u;
          const uint32_t d1 = tmem + kColD1 + j * kN1;
          mbar_wait(&sm.d2_free[0], (u & 1u) ^ 1u, kWaitD2Free);            // E2 has read group 0 of tile u-1
          PROF(2);
#pragma unroll
          for (int c = 0; c < kChunks; ++c, ++g) {
            const uint32_t s = g & 1u;
            if (M6A_D2_SPLIT && c == kGroup1Chunk) mbar_wait(&sm.d2_free[1], (u & 1u) ^ 1u, kWaitD2Free);   // ... and group 1
            mbar_wait(&sm.a_full[s], (g >> 1) & 1u, kWaitAFull);
            PROF(3);
            fence_after();
            const uint32_t a_lo = tmem + kColALo + s * kChunk;
            // two accumulator groups (K-steps of chunks 0..2 | 3..4): fewer truncating accumulations per accumulator
            const uint32_t d2 = tmem + kColD2 + (c >= kGroup1Chunk ? 2 * kN2 : 0);
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < ((M6A_ABL & 4) ? 0 : kChunk / 8); ++ks) {
                const int kstep = c * (kChunk / 8) + ks;
                const uint64_t bw = dw2 + static_cast<uint64_t>(kstep) * (kStepW2 >> 4);
                if (ks == 0 && (c == 0 || c == kGroup1Chunk)) mma_ts(d2, d1 + kstep * 8, bw, idesc64, 0u);   // [main | corr] = A_hi . [W2_hi ; W2_lo]^T
                else mma_ts_acc(d2, d1 + kstep * 8, bw, idesc64);
                mma_ts_acc(d2 + kN2, a_lo + ks * 8, bw, idesc32);            // corr += A_lo . W2_hi^T
              }
              mma_commit(&sm.a_free[s]);
              if (M6A_D2_SPLIT && c == kGroup1Chunk - 1) mma_commit(&sm.d2_full[0][j]);
              if (c == kChunks - 1) mma_commit(&sm.d2_full[1][j]);
            }
            __syncwarp();
            PROF(4);
          }
        }
        if (stop) break;
      }
      PROF_STORE(10, lane == 0);
    }
  } else if (warp >= kE1Warp0 && warp < kE1Warp0 + kRoleWarps) {
    // ======================================== E1: relu + hi/lo split of D1, chunk by chunk ================================
    const int et = tid - kE1Warp0 * 32;                         // 0..255
    const int half = et >> 7;                                   // which 16 columns of every 32-column chunk
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t g = 0;
    PROF_DECL;
    for (uint32_t t = 0;; ++t) {
      const uint32_t b = t & 1u;
      // (not x_full: the staging warps may run two phases of it ahead of this role, and a parity wait cannot lag by two)
      role_wait(&sm.l1_done[b], (t >> 1) & 1u, kWaitL1Done, warp == kE1Warp0, lane, kBarE1, kRoleThreads);
      if (sm.x_ctrl[t % kCtrlRing] == 0) break;                  // the MMA issuer arrives on l1_done for the stop tile too
      fence_after();
      PROF(0);
      const uint32_t d1 = tmem + kColD1 + b * kN1 + half * kHalfCols + lane_base;
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c, ++g) {
        const uint32_t s = g & 1u;
        uint32_t v[kHalfCols], l[kHalfCols];
        if (!(M6A_ABL & 2)) {
          tmem_ld16(d1 + c * kChunk, v);
          wait_ld();
        }
        PROF(1);
        if (!(M6A_ABL & 2)) {
#pragma unroll
          for (int i = 0; i < kHalfCols; ++i) {
            const float h = fmaxf(__uint_as_float(v[i]), 0.0f);
            const float hi = rn_tf32(h);
            v[i] = __float_as_uint(hi);
            l[i] = __float_as_uint(h - hi);
          }
        }
        PROF(2);
        role_wait(&sm.a_free[s], ((g >> 1) & 1u) ^ 1u, kWaitAFree, warp == kE1Warp0, lane, kBarE1, kRoleThreads);   // the MMAs of chunk g-2 have read this A_lo slot
        fence_after();
        PROF(3);
        if (!(M6A_ABL & 2)) {
          tmem_st16(d1 + c * kChunk, v);                                              // A_hi of Linear-2, in place
          tmem_st16(tmem + kColALo + s * kChunk + half * kHalfCols + lane_base, l);   // A_lo
          wait_st();
        }
        fence_before();
        mbar_arrive(&sm.a_full[s]);
        PROF(4);
      }
    }
    PROF_STORE(30, et == 0);
  } else if (warp >= kSeWarp0 && warp < kSeWarp0 + kRoleWarps) {
    // ======================================== staging of X + E2 (sigmoid, outputs) =========================================
    // Two groups of 4 warps (one warp per TMEM lane quadrant, thread = row of the tile) take the tiles alternately: group 0
    // the even ones, group 1 the odd ones -- X[b], x_full[b] and every second use of D2 belong to group b.  The per-tile work of
    // this role is a serial chain (TMEM read -> sigmoid -> outputs -> staging -> prefetch) about twice as long as the tensor
    // pipe needs for a tile; two groups hide it.  Tiles are dealt statically (tile = blockIdx.x + k * gridDim.x), so both
    // groups walk the same slab sequence on their own; the group that meets a slab first (the parity of the slab's first MMA
    // tile) opens it and the other waits for hdr_ready.
    const int grp = (warp - kSeWarp0) >> 2;                     // 0 / 1
    const int row = ((warp & 3) << 5) | lane;                   // row of the tile = TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int emb_dim = image->emb_dim, n_kmer = image->n_kmer;
    const int bar_grp = kBarSe + grp;                           // named barrier of this group (128 threads)

    // ---- slab / tile generator (every thread of both groups runs the same sequence) ------------------------------------
    long long tile = static_cast<long long>(blockIdx.x) - static_cast<long long>(gridDim.x);
    long long tile_s1 = 0, s0 = 0, r0 = 0, s_next = 0;
    int ns = 0, nr = 0, base = 0, slot = -1;
    uint32_t n_slabs = 0;               // slabs met so far (slot = n % kSlots)
    uint32_t t_next = 0;                // index of the next MMA tile of this CTA
    bool exhausted = false, have_tile = false;

    // the slab [s0, s0 + ns) becomes current; `owner`: this group writes its header, the other waits for it
    auto open_slab = [&](bool owner, bool stop_marker) {
      slot = static_cast<int>(n_slabs % kSlots);
      const uint32_t use = n_slabs / kSlots;
      ++n_slabs;
      if (owner) {
        mbar_wait(&sm.slab_empty[slot], (use & 1u) ^ 1u, kWaitSlabEmpty);
        if (stop_marker) {
          if (row == 0) sm.meta[slot].stop = 1;
          return;
        }
        if (row <= ns) sm.roff[slot][row] = static_cast<int>(a.read_off[s0 + row] - r0);
        if (row < ns) {
          sm.cnt[slot][row] = 0;
#pragma unroll
          for (int t = 0; t < kKmerPos; ++t) {
            int k = a.kmer_idx != nullptr ? a.kmer_idx[(s0 + row) * kKmerPos + t] : 0;
            sm.kid[slot][row][t] = min(max(k, 0), n_kmer - 1);
          }
        }
        // most reads of a site -> layout of the slab's q table: site-interleaved when it fits (lanes = sites pooling)
        int my_n = row < ns ? static_cast<int>(a.read_off[s0 + row + 1] - a.read_off[s0 + row]) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_n = max(my_n, __shfl_xor_sync(0xffffffffu, my_n, o));
        if (lane == 0) sm.maxn_scratch[grp][warp & 3] = my_n;
        named_bar_sync(bar_grp, kTileM);
        if (row == 0) {
          SlabMeta m;
          m.s0 = s0; m.r0 = r0; m.ns = ns; m.nr = nr; m.stop = 0; m.pad = 0;
          m.max_n = max(max(sm.maxn_scratch[grp][0], sm.maxn_scratch[grp][1]), max(sm.maxn_scratch[grp][2], sm.maxn_scratch[grp][3]));
          const int n_groups = (ns + 31) >> 5;
          const bool lanes_ok = M6A_TC_LANES_SITES && n_blocks <= kLaneBlocksMax && m.max_n >= 1 &&
                                m.max_n <= static_cast<int>(kPairedMaxReads) && n_groups * 32 * m.max_n <= kQCapT;
          m.mode = lanes_ok ? 1 : (nr <= kQCapT ? 0 : -1);
          sm.meta[slot] = m;
        }
        named_bar_sync(bar_grp, kTileM);                        // this group reads the header right away
        mbar_arrive(&sm.hdr_ready[slot]);
      } else if (!stop_marker) {
        mbar_wait(&sm.hdr_ready[slot], use & 1u, kWaitHdr);
      }
    };

    // next MMA tile of the CTA; `mine`: this group stages it (otherwise only the generator state advances)
    auto next_tile = [&]() -> TileInfo {
      TileInfo ti;
      ti.exists = false; ti.valid = false; ti.last = false; ti.grow = 0; ti.slot = 0; ti.lr = 0; ti.site_l = 0;
      const bool owner = static_cast<int>(t_next & 1u) == grp;   // whoever stages the next tile opens what comes before it
      while (!exhausted && base >= nr) {                          // current slab exhausted (initially nr = 0)
        if (!have_tile || s_next >= tile_s1) {                    // next tile of this CTA
          tile += gridDim.x;
          if (tile >= a.n_tiles) {
            exhausted = true;
            break;
          }
          s_next = a.tile_bounds[tile];
          tile_s1 = a.tile_bounds[tile + 1];
          have_tile = true;
          if (s_next >= tile_s1) continue;                        // a tile without sites
        }
        s0 = s_next;
        ns = static_cast<int>(min(static_cast<long long>(kSlabSites), tile_s1 - s0));
        s_next = s0 + ns;
        r0 = a.read_off[s0];
        nr = static_cast<int>(a.read_off[s0 + ns] - r0);
        base = 0;
        if (nr == 0) {
          // sites without reads have nothing to encode or pool: the owner writes their outputs (NaN, 0) itself.  They take no
          // slab slot, so every slab in the ring holds at least one MMA tile -- which is what keeps the two groups from
          // waiting on each other (a slab's last tile is always at least three tiles behind the first tile of the slab
          // that reuses its slot).
          if (owner && row < ns) {
            const size_t o = static_cast<size_t>(s0 + row) * a.site_stride;
            a.site_prob[o] = __int_as_float(0x7fc00000);
            a.mod_count[o] = 0;
          }
          continue;
        }
        open_slab(owner, false);
      }
      if (exhausted) return ti;
      ti.exists = true;
      ti.slot = slot;
      ti.lr = base + row;
      ti.valid = ti.lr < nr;
      ti.grow = r0 + ti.lr;
      base += kTileM;
      ti.last = base >= nr;
      ++t_next;
      if (owner && ti.valid) {
        int lo = 0, hi = ns;                                       // site of the row: last s with roff[s] <= lr
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (sm.roff[slot][mid] <= ti.lr) lo = mid; else hi = mid;
        }
        ti.site_l = lo;
      }
      return ti;
    };

    // asynchronous prefetch of this row's 16 inputs into fbuf[grp][0..16)[row] (no register round trip)
    auto prefetch_inputs = [&](const TileInfo& ti) {
      float* dst = &sm.fbuf[grp][0][row];                        // input k lives at dst[k * kTileM]
      if (ti.exists && ti.valid) {
        const float* xr = a.feats + ti.grow * kNSig;
#pragma unroll
        for (int k = 0; k < kNSig; ++k) cp_async4(dst + k * kTileM, xr + k);
#pragma unroll
        for (int k = kNSig; k < kK1 - 1; ++k) dst[k * kTileM] = 0.0f;
        if (emb_dim == 2) {
#pragma unroll
          for (int t = 0; t < kKmerPos; ++t) {
            const float* e = image->emb + 2 * sm.kid[ti.slot][ti.site_l][t];
            cp_async4(dst + (kNSig + 2 * t) * kTileM, e);
            cp_async4(dst + (kNSig + 1 + 2 * t) * kTileM, e + 1);
          }
        } else if (emb_dim == 1) {
#pragma unroll
          for (int t = 0; t < kKmerPos; ++t) cp_async4(dst + (kNSig + t) * kTileM, image->emb + sm.kid[ti.slot][ti.site_l][t]);
        }
        dst[(kK1 - 1) * kTileM] = 1.0f;                          // bias column
      } else {
#pragma unroll
        for (int k = 0; k < kK1; ++k) dst[k * kTileM] = 0.0f;
      }
      cp_async_commit();
    };
    // the 16 prefetched inputs of this row -> RN_tf32 split -> the four k-chunks of X[grp]
    auto stage_x = [&]() {
      const float* src = &sm.fbuf[grp][0][row];
#pragma unroll
      for (int jj = 0; jj < kK1 / 4; ++jj) {
        const float4 f = make_float4(src[(4 * jj) * kTileM], src[(4 * jj + 1) * kTileM], src[(4 * jj + 2) * kTileM],
                                     src[(4 * jj + 3) * kTileM]);
        const float4 h = make_float4(rn_tf32(f.x), rn_tf32(f.y), rn_tf32(f.z), rn_tf32(f.w));
        const float4 l = make_float4(f.x - h.x, f.y - h.y, f.z - h.z, f.w - h.w);
        *reinterpret_cast<float4*>(sm.x[grp][0][jj][row]) = h;
        *reinterpret_cast<float4*>(sm.x[grp][1][jj][row]) = l;     // the hardware truncates lo to TF32
      }
    };
    // advance the generator to this group's next tile (the other group's tile in between only moves the state)
    auto next_mine = [&]() -> TileInfo {
      TileInfo ti = next_tile();
      if (ti.exists && static_cast<int>((t_next - 1u) & 1u) != grp) ti = next_tile();
      return ti;
    };

    // Per iteration (this group's k-th tile, t = 2k + grp): stage X(t) | prefetch tile t+2 | E2 of tile t-2.  The MMA issuer needs
    // X(t) before it needs D2 back (Linear-1 of tile t precedes Linear-2 of tile t-1), and E2 has to wait for Linear-2 of
    // tile t-2 anyway.
    TileInfo ti_s = next_mine();          // tile staged in this iteration
    uint32_t t_s = t_next - 1u;           // its index (meaningful when ti_s.exists)
    TileInfo ti_e;                        // tile whose E2 is pending (staged one iteration ago)
    uint32_t t_e = 0;
    ti_e.exists = false;
    prefetch_inputs(ti_s);
    bool stop_sent = false;
    PROF_DECL;
    for (;;) {
      // ---- (2) stage X(t) (or publish the stop) ---------------------------------------------------------------------------
      TileInfo ti_n;
      ti_n.exists = false;
      uint32_t t_n = 0;
      if (!stop_sent) {
        if (ti_s.exists) {
          if (t_s >= 2) mbar_wait(&sm.l1_done[grp], ((t_s - 2) >> 1) & 1u, kWaitL1DoneSe);   // Linear-1 of tile t-2 has consumed X[grp]
          PROF(0);
          cp_async_wait_all();
          PROF(1);
          if (!(M6A_ABL & 16)) stage_x();
          if (row == 0) sm.x_ctrl[t_s % kCtrlRing] = 1;
          PROF(2);
          if (!(M6A_ABL & 1)) fence_proxy_async();
          PROF(3);
          mbar_arrive(&sm.x_full[grp]);
          PROF(4);
          // ---- (3) the inputs of this group's next tile travel to shared memory meanwhile ----------------------------------
          ti_n = next_mine();
          t_n = t_next - 1u;
          PROF(5);
          if (!(M6A_ABL & 16)) prefetch_inputs(ti_n);
          PROF(6);
        } else {
          // no tile left for this group: the group whose turn the next tile index is tells the MMA issuer to stop
          if (static_cast<int>(t_next & 1u) == grp) {
            // x_full[grp] must not complete a second phase before the MMA issuer has consumed the first one (a parity wait
            // that lags two phases never returns): the stop goes out only after Linear-1 of this group's last tile.
            if (t_next >= 2) mbar_wait(&sm.l1_done[grp], ((t_next - 2) >> 1) & 1u, kWaitL1DoneSe);
            if (row == 0) sm.x_ctrl[t_next % kCtrlRing] = 0;
            mbar_arrive(&sm.x_full[grp]);
          }
          stop_sent = true;
        }
      }
      // ---- (1) E2 of the tile staged one iteration ago: all 32 outputs of this row, 8 at a time ------------------------------
      if (ti_e.exists) {
        const uint32_t d2 = tmem + kColD2 + lane_base;
#if M6A_D2_SPLIT
        // accumulator group 0 (chunks 0..2) is complete two chunks before the tile is: read it out and hand it back early
        mbar_wait(&sm.d2_full[0][grp], (t_e >> 1) & 1u, kWaitD2Full);
        fence_after();
        PROF(7);
        float* g0 = &sm.g0[grp][0][row];                           // (main + corr) of group 0: g0[k * kTileM]
#pragma unroll
        for (int o8 = 0; o8 < ((M6A_ABL & 8) ? 0 : kN2); o8 += 8) {
          uint32_t m0[8], c0[8];
          tmem_ld8(d2 + o8, m0);
          tmem_ld8(d2 + kN2 + o8, c0);
          wait_ld();
#pragma unroll
          for (int k = 0; k < 8; ++k) g0[(o8 + k) * kTileM] = __uint_as_float(m0[k]) + __uint_as_float(c0[k]);
        }
        fence_before();
        mbar_arrive(&sm.d2_free[0]);
        mbar_wait(&sm.d2_full[1][grp], (t_e >> 1) & 1u, kWaitD2Full);
        fence_after();
        float z = sm.b3;
#pragma unroll
        for (int o8 = 0; o8 < ((M6A_ABL & 8) ? 0 : kN2); o8 += 8) {
          uint32_t m1[8], c1[8];
          tmem_ld8(d2 + 2 * kN2 + o8, m1);
          tmem_ld8(d2 + 3 * kN2 + o8, c1);
          wait_ld();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // (main + corr) of each accumulator group, groups added in K order, then the bias: all round-to-nearest float32
            const float h2 = (g0[(o8 + k) * kTileM] + (__uint_as_float(m1[k]) + __uint_as_float(c1[k]))) + sm.b2[o8 + k];
            z = fmaf(sm.w3[o8 + k], fmaxf(h2, 0.0f), z);
          }
        }
        fence_before();
        mbar_arrive(&sm.d2_free[1]);
#else
        mbar_wait(&sm.d2_full[1][grp], (t_e >> 1) & 1u, kWaitD2Full);
        fence_after();
        PROF(7);
        float z = sm.b3;
#pragma unroll
        for (int o8 = 0; o8 < ((M6A_ABL & 8) ? 0 : kN2); o8 += 8) {
          uint32_t m0[8], c0[8], m1[8], c1[8];
          tmem_ld8(d2 + o8, m0);
          tmem_ld8(d2 + kN2 + o8, c0);
          tmem_ld8(d2 + 2 * kN2 + o8, m1);
          tmem_ld8(d2 + 3 * kN2 + o8, c1);
          wait_ld();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float h2 = ((__uint_as_float(m0[k]) + __uint_as_float(c0[k])) + (__uint_as_float(m1[k]) + __uint_as_float(c1[k]))) +
                             sm.b2[o8 + k];
            z = fmaf(sm.w3[o8 + k], fmaxf(h2, 0.0f), z);
          }
        }
        fence_before();
        mbar_arrive(&sm.d2_free[0]);
#endif
        PROF(8);
        const float p = 1.0f / (1.0f + expf(-z));
        if (ti_e.valid) {
          a.read_prob[ti_e.grow] = p;
          {
            const int mode = sm.meta[ti_e.slot].mode;
            if (mode == 1) {            // [group][entry][32 sites]
              const int entry = ti_e.lr - sm.roff[ti_e.slot][ti_e.site_l];
              sm.q[ti_e.slot][((ti_e.site_l >> 5) * sm.meta[ti_e.slot].max_n + entry) * 32 + (ti_e.site_l & 31)] = 1.0f - p;
            } else if (mode == 0) {
              sm.q[ti_e.slot][ti_e.lr] = 1.0f - p;
            }
          }
          if (p >= a.read_threshold) atomicAdd(&sm.cnt[ti_e.slot][ti_e.site_l], 1);
        }
        if (ti_e.last) mbar_arrive(&sm.slab_full[ti_e.slot]);
        PROF(9);
      }
      // ---- (4) rotate -------------------------------------------------------------------------------------------------------
      ti_e = ti_s;
      t_e = t_s;
      ti_s = ti_n;
      t_s = t_n;
      if (stop_sent && !ti_e.exists) break;
    }
    PROF_STORE(0, grp == 0 && row == 0);
    // the stop marker for the MC warps comes from the group whose turn it is
    {
      const bool owner = static_cast<int>(t_next & 1u) == grp;
      open_slab(owner, true);
      if (owner) mbar_arrive(&sm.slab_full[slot]);
    }
  } else {
    // ======================================== Monte-Carlo pooling ============================================================
    // A warp takes whole sites of a finished slab (w, w + 15, ...): the blocks of 32 * ipl iterations of one site are the
    // interleaved chains, the block sums are added in block order (the summation order of m6a_kernel.cu: results are
    // bit-identical functions of the per-read probabilities), and the warp writes the site's outputs itself.  No barrier
    // between slabs: a warp moves on as soon as its sites are done; the last one releases the slot.
#if M6A_TC_LAYOUT == 0
    const int mcw = warp == 0 ? 0 : (warp < 4 ? warp - 1 : warp - 17);     // 0, 2, 3, 20..31 -> 0..14
#else
    const int mcw = warp < 12 ? warp : warp - 16;                          // 0..11, 28..30 -> 0..14
#endif
    const float n_iters_f = static_cast<float>(a.n_iters);
    PROF_DECL;
    for (uint32_t n = 0;; ++n) {
      const int slot = static_cast<int>(n % kSlots);
      mbar_wait(&sm.slab_full[slot], (n / kSlots) & 1u, kWaitSlabFull);
      PROF(0);
      const SlabMeta m = sm.meta[slot];
      if (m.stop) break;
      const int ns = m.ns;
      const bool q_in_smem = m.mode == 0;
      const int* roff = sm.roff[slot];
      const float* q = sm.q[slot];
#if M6A_TC_LANES_SITES
      if (m.mode == 1 && !(M6A_ABL & 32)) {
        // ---- lanes = sites: hardware lane L pools site 32 g + L; an item = (group g, block b, octet k) = the four lane
        // streams {k, k+16, k+8, k+24} of block b, i.e. the partial sum b_k of the butterfly sum of m6a_mc.cuh, so the final
        // result is the same float32 expression as everywhere else.  Every lane of a warp reads its own column of the
        // site-interleaved q table: one shared-memory wavefront per load whatever the indices are.
        const int n_groups = (ns + 31) >> 5;
        const int items = n_groups * n_blocks * kOctets;
        for (int item = mcw; item < items; item += kMcWarps) {
          const int g = item / (n_blocks * kOctets), rem = item - g * (n_blocks * kOctets);
          const int b = rem / kOctets, k = rem - b * kOctets;
          const int sl = 32 * g + lane;
          const bool live = sl < ns;
          const int nreads = live ? roff[sl + 1] - roff[sl] : 0;
          const uint32_t n = nreads > 0 ? static_cast<uint32_t>(nreads) : 1u;        // idle lanes read entry 0 of their column
          const unsigned long long site_id = static_cast<unsigned long long>(a.site_id_base + m.s0 + (live ? sl : 0));
          const uint32_t qaddr = mc_smem_u32(q + static_cast<size_t>(g) * m.max_n * 32 + lane);
          float v[4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {                // (k, k+16) then (k+8, k+24): two chains at a time
            const int vl0 = k + 8 * h, vl1 = vl0 + 16;
            Mwc64x g0, g1;
            g0.seed(static_cast<uint32_t>(vl0), static_cast<uint32_t>(b), site_id, a.seed);
            g1.seed(static_cast<uint32_t>(vl1), static_cast<uint32_t>(b), site_id, a.seed);
            const long long it0 = static_cast<long long>(b) * ipl * 32;
            const long long l0 = (static_cast<long long>(a.n_iters) - (it0 + vl0) + 31) / 32;
            const long long l1 = (static_cast<long long>(a.n_iters) - (it0 + vl1) + 31) / 32;
            const int r0_ = static_cast<int>(l0 < 0 ? 0 : (l0 > ipl ? ipl : l0)), r1_ = static_cast<int>(l1 < 0 ? 0 : (l1 > ipl ? ipl : l1));
            float s0_ = 0.0f, s1_ = 0.0f;
            const int rc = min(r0_, r1_);
            for (int r = 0; r < rc; ++r) {
              float p0 = 1.0f, p1 = 1.0f;
#pragma unroll
              for (int s = 0; s < NS / 2; ++s) {
                uint32_t i1, i2, j1, j2;
                g0.next_pair(n, i1, i2);
                g1.next_pair(n, j1, j2);
                p0 *= lds_f32(qaddr + (i1 << 7));
                p1 *= lds_f32(qaddr + (j1 << 7));
                p0 *= lds_f32(qaddr + (i2 << 7));
                p1 *= lds_f32(qaddr + (j2 << 7));
              }
              s0_ += 1.0f - p0;
              s1_ += 1.0f - p1;
            }
            for (int r = rc; r < r0_; ++r) {
              float p0 = 1.0f;
#pragma unroll
              for (int s = 0; s < NS / 2; ++s) {
                uint32_t i1, i2;
                g0.next_pair(n, i1, i2);
                p0 *= lds_f32(qaddr + (i1 << 7));
                p0 *= lds_f32(qaddr + (i2 << 7));
              }
              s0_ += 1.0f - p0;
            }
            for (int r = rc; r < r1_; ++r) {
              float p1 = 1.0f;
#pragma unroll
              for (int s = 0; s < NS / 2; ++s) {
                uint32_t j1, j2;
                g1.next_pair(n, j1, j2);
                p1 *= lds_f32(qaddr + (j1 << 7));
                p1 *= lds_f32(qaddr + (j2 << 7));
              }
              s1_ += 1.0f - p1;
            }
            v[2 * h] = s0_;
            v[2 * h + 1] = s1_;
          }
          sm.part[slot][g][b][k][lane] = (v[0] + v[1]) + (v[2] + v[3]);       // b_k = (v_k + v_k+16) + (v_k+8 + v_k+24)
        }
        PROF(1);
        __syncwarp();
        bool last = false;
        if (lane == 0) {
          __threadfence_block();
          last = atomicAdd(&sm.done[slot], 1) == kMcWarps - 1;
        }
        last = __shfl_sync(0xffffffffu, last ? 1 : 0, 0) != 0;
        if (last) {                      // every item of the slab is in: finish the butterfly sums, write the sites
          __threadfence_block();
          for (int g = 0; g < n_groups; ++g) {
            const int sl = 32 * g + lane;
            if (sl < ns) {
              const int nreads = roff[sl + 1] - roff[sl];
              float total = 0.0f;
              for (int b = 0; b < n_blocks; ++b) {
                const float (&bk)[kOctets][32] = sm.part[slot][g][b];
                const float c0 = bk[0][lane] + bk[4][lane], c1 = bk[1][lane] + bk[5][lane];
                const float c2 = bk[2][lane] + bk[6][lane], c3 = bk[3][lane] + bk[7][lane];
                total += (c0 + c2) + (c1 + c3);
              }
              const size_t o = static_cast<size_t>(m.s0 + sl) * a.site_stride;
              a.site_prob[o] = nreads > 0 ? total / n_iters_f : __int_as_float(0x7fc00000);
              a.mod_count[o] = sm.cnt[slot][sl];
            }
          }
          __syncwarp();
          if (lane == 0) {
            sm.done[slot] = 0;
            mbar_arrive(&sm.slab_empty[slot]);
          }
        }
        PROF(3);
        continue;
      }
#endif

      auto lane_rounds = [&](int blk_) {
        const long long it0 = static_cast<long long>(blk_) * ipl * 32 + lane;
        const long long left = (static_cast<long long>(a.n_iters) - it0 + 31) / 32;
        return static_cast<int>(left < 0 ? 0 : (left > ipl ? ipl : left));
      };
      for (int sl = mcw; sl < ((M6A_ABL & 32) ? 0 : ns); sl += kMcWarps) {
        const int nreads = roff[sl + 1] - roff[sl];
        const unsigned long long site_id = static_cast<unsigned long long>(a.site_id_base + m.s0 + sl);
        float total = 0.0f;
        if (nreads > 0) {
          const bool paired = nreads <= static_cast<int>(kPairedMaxReads);
          for (int b0 = 0; b0 < n_blocks; b0 += kMcChains) {
            if (q_in_smem && b0 + kMcChains <= n_blocks) {      // warp-uniform: kMcChains blocks of this site at once
              uint32_t qa[kMcChains], nu[kMcChains];
              Mwc64x gen[kMcChains];
              int rounds[kMcChains];
              float v[kMcChains];
#pragma unroll
              for (int i = 0; i < kMcChains; ++i) {
                qa[i] = mc_smem_u32(q + roff[sl]);
                nu[i] = static_cast<uint32_t>(nreads);
                rounds[i] = lane_rounds(b0 + i);
                gen[i].seed(static_cast<uint32_t>(lane), static_cast<uint32_t>(b0 + i), site_id, a.seed);
                v[i] = 0.0f;
              }
              if (paired) mc_rounds_xn<NS, true, kMcChains>(qa, nu, gen, rounds, v);
              else mc_rounds_xn<NS, false, kMcChains>(qa, nu, gen, rounds, v);
#pragma unroll
              for (int i = 0; i < kMcChains; ++i) total += warp_butterfly_sum(v[i]);
            } else {
#pragma unroll 1
              for (int blk = b0; blk < n_blocks && blk < b0 + kMcChains; ++blk) {
                const int rounds = lane_rounds(blk);
                Mwc64x gen;
                gen.seed(static_cast<uint32_t>(lane), static_cast<uint32_t>(blk), site_id, a.seed);
                float v;
                if (q_in_smem) v = mc_lane_smem<NS>(q + roff[sl], static_cast<uint32_t>(nreads), gen, rounds);
                else v = mc_lane_generic(a.read_prob + m.r0 + roff[sl], true, static_cast<uint32_t>(nreads), gen, rounds, NS, nullptr, 0);
                total += warp_butterfly_sum(v);
              }
            }
          }
        }
        if (lane == 0) {
          const size_t o = static_cast<size_t>(m.s0 + sl) * a.site_stride;
          a.site_prob[o] = nreads > 0 ? total / n_iters_f : __int_as_float(0x7fc00000);
          a.mod_count[o] = sm.cnt[slot][sl];
        }
      }
      PROF(1);
      // every warp reports once per slab (also with no site of its own: its slab_full phase must not fall two behind)
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        if (atomicAdd(&sm.done[slot], 1) == kMcWarps - 1) {
          sm.done[slot] = 0;
          mbar_arrive(&sm.slab_empty[slot]);
        }
      }
      PROF(3);
    }
    PROF_STORE(20, mcw == 0 && lane == 0);
  }

  // ---- teardown -----------------------------------------------------------------------------------------------------------
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, kTmemCols);
}

// ---- host side -------------------------------------------------------------------------------------------------------------
static bool g_tc_attr_done[64];
static int g_tc_smem_set[64];

cudaError_t set_trap_record(int* mapped_device_ptr) {
  return cudaMemcpyToSymbol(m6a::tc::g_trap_record, &mapped_device_ptr, sizeof(int*));
}

}  // namespace tcx

cudaError_t launch_mil_infer_tc(const KernelArgs& a, const tcx::WeightImageTc* d_image, int n_sms, cudaStream_t stream,
                                LaunchInfo* info) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  const int smem = static_cast<int>(tcx::tc_smem_bytes(a.n_blocks));
  auto kern = tcx::mil_infer_tc_kernel<20>;
  if (!tcx::g_tc_attr_done[dev] || tcx::g_tc_smem_set[dev] < smem) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    tcx::g_tc_attr_done[dev] = true;
    tcx::g_tc_smem_set[dev] = smem;
  }
  long long grid = n_sms;
  if (grid > a.n_tiles) grid = a.n_tiles;
  if (grid < 1) grid = 1;
  if (info) {
    info->grid = static_cast<int>(grid);
    info->block = tcx::kTcThreads;
    info->smem_bytes = smem;
    info->tile_reads = a.tile_reads;
  }
  kern<<<static_cast<unsigned>(grid), tcx::kTcThreads, smem, stream>>>(a, d_image);
  return cudaGetLastError();
}

cudaError_t tc_set_trap_record(int* mapped_device_ptr) { return tcx::set_trap_record(mapped_device_ptr); }

cudaError_t tc_read_profile(unsigned long long* out32) {
#if M6A_TC_PROFILE
  return cudaMemcpyFromSymbol(out32, tcx::g_prof, 40 * sizeof(unsigned long long));
#else
  for (int i = 0; i < 40; ++i) out32[i] = 0;
  return cudaSuccess;
#endif
}

}  // namespace m6a
