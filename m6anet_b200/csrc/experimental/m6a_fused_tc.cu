// EXPERIMENTAL fused MIL-inference kernel with the tensor-core read encoder: phase A of the product kernel
// (m6anet_b200/csrc/m6a_kernel.cu) replaced by tc_encode_tile (m6a_tc_device.cuh), phase B (Monte-Carlo noisy-OR on the
// Philox-seeded MWC64X lane streams, m6a_mc.cuh) unchanged, so site probabilities stay bit-identical functions of the
// per-read probabilities and of (seed, global site id, n_iters).  Same status as m6a_encoder_tc.cu: cross-compiled only,
// never run on a GPU, built into the experimental library that the product does not load.
//
// CTA = 256 threads (8 warps: two per TMEM lane quadrant for the relu/split chunks, all eight for phase B), 2 CTAs per SM
// (256 TMEM columns and ~110 KB of shared memory each).  Per slice of <= 64 consecutive sites (read-balanced tiles as in
// the product kernel): header -> for every 128 reads: gather [x | emb | 1] rows, tensor-core encoder, p -> read_prob /
// q = 1 - p in shared memory / threshold count -> phase B -> site_prob, mod_count.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#include "../../../include/m6anet_b200.h"
#include "../m6a_kernel.h"
#include "../m6a_mc.cuh"
#include "m6a_encoder_tc.h"
#include "m6a_tc_device.cuh"

namespace m6a {
namespace tc {

constexpr int kFusedThreads = 256;
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kFusedSites = 64;      // sites per slice
constexpr int kFusedQCap = 4096;     // q = 1 - p entries kept in shared memory per slice

struct alignas(128) FusedSmem {
  TcSmem tc;
  float q[kFusedQCap];
  float partial[kFusedSites][kMaxBlocks];
  int roff[kFusedSites + 1];
  int cnt[kFusedSites];
  int kid[kFusedSites][kKmerPos];
};

struct FusedArgs {
  KernelArgs k;                  // the product kernel's argument block (model / sample_idx / feats_tma_ok unused)
  const WeightImageTc* image;
  int n_kmer, emb_dim;
};

template <int NS>
__global__ void __launch_bounds__(kFusedThreads, 2)
mil_infer_tc_kernel(const FusedArgs fa) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
  const KernelArgs& a = fa.k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  TcState st;
  tc_setup(sm.tc, st, fa.image, tid, kFusedThreads);

  const int n_blocks = a.n_blocks, ipl = a.iters_per_lane;
  const float n_iters_f = static_cast<float>(a.n_iters);

  for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
   const long long tile_s0 = a.tile_bounds[tile], tile_s1 = a.tile_bounds[tile + 1];
   for (long long s0 = tile_s0; s0 < tile_s1; s0 += kFusedSites) {
    const int ns = static_cast<int>(min(static_cast<long long>(kFusedSites), tile_s1 - s0));
    const long long r0 = a.read_off[s0];
    // ---- slice header -------------------------------------------------------------------------------------------------
    if (tid <= ns) sm.roff[tid] = static_cast<int>(a.read_off[s0 + tid] - r0);
    if (tid < ns) {
      sm.cnt[tid] = 0;
#pragma unroll
      for (int t = 0; t < kKmerPos; ++t) {
        int k = a.kmer_idx != nullptr ? a.kmer_idx[(s0 + tid) * kKmerPos + t] : 0;
        sm.kid[tid][t] = min(max(k, 0), fa.n_kmer - 1);
      }
    }
    __syncthreads();
    const int nr = sm.roff[ns];
    const bool q_in_smem = nr <= kFusedQCap;

    // ---- phase A: 128 reads per pass through the tensor-core encoder ---------------------------------------------------
    for (int base = 0; base < nr; base += kTileM) {
      const int lr = base + tid;                            // slice-local read (threads 0..127 own one row each)
      const bool valid = tid < kTileM && lr < nr;
      int site_l = 0;
      if (tid < kTileM) {
        float in[kK1];
#pragma unroll
        for (int k = 0; k < kK1; ++k) in[k] = 0.0f;
        if (valid) {
          int lo = 0, hi = ns;                              // site of the read: last s with roff[s] <= lr
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (sm.roff[mid] <= lr) lo = mid; else hi = mid;
          }
          site_l = lo;
          const float* xr = a.feats + (r0 + lr) * kNSig;
#pragma unroll
          for (int k = 0; k < kNSig; ++k) in[k] = xr[k];
          if (fa.emb_dim > 0) {
#pragma unroll
            for (int t = 0; t < kKmerPos; ++t) {
              const int kid = sm.kid[lo][t];
              if (fa.emb_dim == 2) {
                in[kNSig + 2 * t] = __ldg(fa.image->emb + kid * 2);
                in[kNSig + 2 * t + 1] = __ldg(fa.image->emb + kid * 2 + 1);
              } else {
                in[kNSig + t] = __ldg(fa.image->emb + kid);
              }
            }
          }
          in[kK1 - 1] = 1.0f;                               // bias column
        }
        tc_stage_row(sm.tc, tid, in);
      }
      const float p = tc_encode_tile<2>(sm.tc, st, tid);
      if (valid) {
        a.read_prob[r0 + lr] = p;
        if (q_in_smem) sm.q[lr] = 1.0f - p;
        if (p >= a.read_threshold) atomicAdd(&sm.cnt[site_l], 1);
      }
    }
    __syncthreads();   // q, cnt and read_prob of the whole slice are visible

    // ---- phase B: Monte-Carlo noisy-OR (as m6a_kernel.cu: items (site, block) dealt to the warps two at a time) --------
    {
      const int items = ns * n_blocks;
      auto advance = [&](int& sl_, int& blk_) {
        blk_ += kFusedWarps;
        while (blk_ >= n_blocks) { blk_ -= n_blocks; ++sl_; }
      };
      auto lane_rounds = [&](int blk_, long long& it0) {
        it0 = static_cast<long long>(blk_) * ipl * 32 + lane;
        const long long left = (static_cast<long long>(a.n_iters) - it0 + 31) / 32;
        return static_cast<int>(left < 0 ? 0 : (left > ipl ? ipl : left));
      };
      auto run_single = [&](int sl_, int blk_) {
        const int n = sm.roff[sl_ + 1] - sm.roff[sl_];
        float v = 0.0f;
        if (n > 0) {
          long long it0;
          const int rounds = lane_rounds(blk_, it0);
          Mwc64x g;
          g.seed(static_cast<uint32_t>(lane), static_cast<uint32_t>(blk_),
                 static_cast<unsigned long long>(a.site_id_base + s0 + sl_), a.seed);
          if (NS > 0 && q_in_smem) {
            v = mc_lane_smem<NS>(sm.q + sm.roff[sl_], static_cast<uint32_t>(n), g, rounds);
          } else {
            const float* qbase = q_in_smem ? (sm.q + sm.roff[sl_]) : (a.read_prob + r0 + sm.roff[sl_]);
            v = mc_lane_generic(qbase, !q_in_smem, static_cast<uint32_t>(n), g, rounds, a.n_samples, nullptr, 0);
          }
        }
        v = warp_butterfly_sum(v);
        if (lane == 0) sm.partial[sl_][blk_] = v;
      };
      int sl = 0, blk = warp;
      while (blk >= n_blocks) { blk -= n_blocks; ++sl; }
      for (int item = warp; item < items; item += 2 * kFusedWarps) {
        int sl2 = sl, blk2 = blk;
        advance(sl2, blk2);
        const bool have2 = item + kFusedWarps < items;
        const int na = sm.roff[sl + 1] - sm.roff[sl];
        const int nb = have2 ? sm.roff[sl2 + 1] - sm.roff[sl2] : 0;
        const bool fast2 = NS > 0 && q_in_smem && have2 && na > 0 && nb > 0 &&
                           ((na <= static_cast<int>(kPairedMaxReads)) == (nb <= static_cast<int>(kPairedMaxReads)));
        if (fast2) {
          long long it0a, it0b;
          const int ra = lane_rounds(blk, it0a), rb = lane_rounds(blk2, it0b);
          Mwc64x ga, gb;
          ga.seed(static_cast<uint32_t>(lane), static_cast<uint32_t>(blk),
                  static_cast<unsigned long long>(a.site_id_base + s0 + sl), a.seed);
          gb.seed(static_cast<uint32_t>(lane), static_cast<uint32_t>(blk2),
                  static_cast<unsigned long long>(a.site_id_base + s0 + sl2), a.seed);
          float va, vb;
          mc_lane_smem_x2<NS>(sm.q + sm.roff[sl], static_cast<uint32_t>(na), ga, ra, va, sm.q + sm.roff[sl2],
                              static_cast<uint32_t>(nb), gb, rb, vb);
          va = warp_butterfly_sum(va);
          vb = warp_butterfly_sum(vb);
          if (lane == 0) {
            sm.partial[sl][blk] = va;
            sm.partial[sl2][blk2] = vb;
          }
        } else {
          run_single(sl, blk);
          if (have2) run_single(sl2, blk2);
        }
        sl = sl2;
        blk = blk2;
        advance(sl, blk);
      }
    }
    __syncthreads();

    // ---- finalize ------------------------------------------------------------------------------------------------------
    if (tid < ns) {
      const int n = sm.roff[tid + 1] - sm.roff[tid];
      float s = 0.0f;
      for (int k = 0; k < n_blocks; ++k) s += sm.partial[tid][k];
      a.site_prob[s0 + tid] = n > 0 ? s / n_iters_f : __int_as_float(0x7fc00000);
      a.mod_count[s0 + tid] = sm.cnt[tid];
    }
    __syncthreads();
   }
  }
  tc_teardown(sm.tc, st, tid);
}

// tile t starts at the first site whose first feature row is >= t * tile_reads (as tile_bounds_kernel of the product)
__global__ void tc_tile_bounds_kernel(const int64_t* __restrict__ read_off, long long n_sites, long long n_tiles, int tile_reads,
                                      long long* __restrict__ tile_bounds) {
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t > n_tiles) return;
  if (t == n_tiles) {
    tile_bounds[t] = n_sites;
    return;
  }
  const long long target = t * tile_reads;
  long long lo = 0, hi = n_sites;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (read_off[mid] < target) lo = mid + 1; else hi = mid;
  }
  tile_bounds[t] = lo;
}

}  // namespace tc
}  // namespace m6a

// Whole hot path with the tensor-core encoder: arguments as m6a_mil_infer_f32 (include/m6anet_b200.h) without explicit
// indices; workspace >= (total_reads / 64 + 2) * 8 bytes.  EXPERIMENTAL.
extern "C" int m6a_tc_mil_infer_f32(const m6a_tc_encoder* h, const float* feats, const int64_t* read_off,
                                    const int32_t* kmer_idx, int64_t n_sites, int64_t total_reads, int64_t site_id_base,
                                    int32_t n_samples, int32_t n_iters, uint64_t seed, float read_threshold,
                                    float* read_prob, float* site_prob, int32_t* mod_count, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
  using namespace m6a;
  if (!h || n_sites < 0 || total_reads < 0) return M6A_EINVAL;
  if (n_samples < 1 || n_samples > kMaxSamples || n_iters < 1) return M6A_EINVAL;
  if (n_sites == 0) return M6A_OK;
  if (!read_off || !site_prob || !mod_count || !workspace) return M6A_EINVAL;
  if (total_reads > 0 && (!feats || !read_prob)) return M6A_EINVAL;
  if (h->emb_dim > 0 && !kmer_idx) return M6A_EINVAL;
  tc::FusedArgs fa;
  memset(&fa, 0, sizeof(fa));
  KernelArgs& a = fa.k;
  a.feats = feats;
  a.read_off = read_off;
  a.kmer_idx = kmer_idx;
  a.read_prob = read_prob;
  a.site_prob = site_prob;
  a.mod_count = mod_count;
  a.n_sites = n_sites;
  // rows per tile: a multiple of the site depth close to 1024 (8 encoder passes of 128 reads) when the depth is constant
  long long T = 1024;
  if (total_reads % n_sites == 0) {
    const long long depth = total_reads / n_sites;
    if (depth >= 1 && depth <= T) T = (T / depth) * depth;
  }
  a.tile_reads = static_cast<int>(std::max<long long>(64, T));
  a.n_tiles = total_reads / a.tile_reads + 1;
  if (workspace_bytes < static_cast<int64_t>((a.n_tiles + 1) * sizeof(long long))) return M6A_EINVAL;
  a.tile_bounds = static_cast<const long long*>(workspace);
  a.site_id_base = site_id_base;
  a.seed = seed;
  a.n_samples = n_samples;
  a.n_iters = n_iters;
  block_layout(n_iters, &a.iters_per_lane, &a.n_blocks);
  a.read_threshold = read_threshold;
  fa.image = static_cast<const tc::WeightImageTc*>(h->d_image);
  fa.n_kmer = h->n_kmer;
  fa.emb_dim = h->emb_dim;

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long nb = a.n_tiles + 1;
  tc::tc_tile_bounds_kernel<<<static_cast<unsigned>((nb + 255) / 256), 256, 0, st>>>(read_off, n_sites, a.n_tiles, a.tile_reads,
                                                                                   static_cast<long long*>(workspace));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  const int smem = static_cast<int>(sizeof(tc::FusedSmem));
  long long grid = std::min<long long>(static_cast<long long>(h->n_sms) * 2, a.n_tiles);
  if (n_samples == 20) {
    e = cudaFuncSetAttribute(tc::mil_infer_tc_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    tc::mil_infer_tc_kernel<20><<<static_cast<unsigned>(grid), tc::kFusedThreads, smem, st>>>(fa);
  } else {
    e = cudaFuncSetAttribute(tc::mil_infer_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    tc::mil_infer_tc_kernel<0><<<static_cast<unsigned>(grid), tc::kFusedThreads, smem, st>>>(fa);
  }
  e = cudaGetLastError();
  return e == cudaSuccess ? M6A_OK : static_cast<int>(e);
}
