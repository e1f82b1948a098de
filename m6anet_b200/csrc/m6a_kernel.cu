// Fused MIL-inference kernel for sm_100a: read encoder -> q = 1-p in shared memory ->
// Monte-Carlo noisy-OR mean -> site probability, mod_count, per-read probability.
//
// Replaces, per tile of consecutive sites, the reference's
//   model.get_read_representation + pooling_filter.probability_layer   utils/inference_utils.py:35-37
//   group_results / mod_ratio                                          utils/inference_utils.py:48-53
//   calculate_site_proba (_calculate_site_proba)                       utils/inference_utils.py:54,74-104
//
// Work decomposition (one persistent CTA loops over tiles; a tile = the consecutive sites whose first feature row
// lies in [t*T, (t+1)*T), T = tile_reads, so every tile carries about T rows however uneven the sites are;
// the boundaries come from a prepass, tile_bounds_kernel):
//   stage   the tile's feature rows are one contiguous byte range of `feats` -> one cp.async.bulk
//           (TMA, mbarrier completion) per chunk of kChunkReads rows into shared memory
//   phase A thread-per-read (kReadsPerThread reads per thread), hidden units two at a time with
//           packed FFMA2 (B200 issues FFMA2 at half the FFMA rate but each does two FMAs, which
//           frees issue slots for the integer work of phase B running in the co-resident CTA):
//             (h_j0,h_j1) = relu(c_site[j0:j1] + sum_k (w1[j0,k],w1[j1,k]) * x_k)     9 FFMA2
//             acc[0:32]  += w2[:,j0] * h_j0 ; acc[0:32] += w2[:,j1] * h_j1            32 FFMA2
//           the weight image is a __grid_constant__ kernel parameter: LDCU.64 into uniform registers, FFMA2 with a
//           uniform-register operand -- no shared-memory traffic, no vector registers for weights (m6a_layout.h);
//           the k-mer embedding part of Linear-1 is a per-site constant c_site (m6a_layout.h)
//   phase B warp-per-(site, block of 32*ipl iterations): every lane owns a Philox-seeded MWC64X
//           stream (m6a_rng.cuh) and runs its ipl iterations: 20 x (draw, index, LDS q[idx], FMUL),
//           two indices per 32-bit word for sites with <= 256 reads
//   final   butterfly sum per block, ordered sum of block partials / n_iters -> site_prob
// The summation order and the index streams are functions of (seed, site id, n_iters) only -- not of
// tiling, grid or GPU count -- so a site's result is bit-identical however the sites are sharded.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "m6a_kernel.h"
#include "m6a_mc.cuh"
#include "m6a_rng.cuh"

namespace m6a {

struct Smem {
#if !M6A_WEIGHTS_CONST
  WeightImage w;                                    // staged once per CTA by one bulk copy
#endif
  alignas(16) float feat[kChunkReads * kNSig + 8];  // + unaligned head (<=3 floats) + tail round-up
  alignas(16) float csite[kSitesPerTileMax][kCStride];
  float q[kQCap];
  float partial[kSitesPerTileMax][kMaxBlocks];
  int roff[kSitesPerTileMax + 1];
  int cnt[kSitesPerTileMax];
  int kid[kSitesPerTileMax][kKmerPos];
  alignas(8) unsigned long long bar_w;
  alignas(8) unsigned long long bar_f;
};

size_t smem_bytes() { return sizeof(Smem); }

// ---- mbarrier / bulk-copy (TMA 1-D) primitives -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

__device__ __forceinline__ float2 ffma2(float2 a, float b, float2 c) {
  return __ffma2_rn(a, make_float2(b, b), c);   // SASS: FFMA2 Rd, Ra.F32x2.HI_LO, Rb.F32, Rc.F32x2.HI_LO
}

// validate()-style bags: `rounds` passes of ONE lane, one bag of k reads per pass, pooled like the model's pooling block
// (reference model_blocks/pooling_blocks.py:96-98 mean, :127-129 noisy-OR, :158-160 max) on the per-read probabilities
// this CTA wrote in phase A.  Bag source: explicit indices, Floyd draws without replacement (m6a_rng.cuh), or the
// inference stream.  Writes each pass to bag_out[32 * pass] (row (site) of bag_prob [n_sites, n_iters], iterations of a
// lane are 32 apart) and returns the lane's sum over its passes.
template <int NS>
__device__ __forceinline__ float bag_rounds(const float* pbase, uint32_t n, Mwc64x& g, int rounds, int k_rt,
                                            const BagArgs& bags, const uint16_t* __restrict__ explicit_idx,
                                            size_t explicit_round_stride, float* bag_out) {
  constexpr int kCap = NS > 0 ? NS : kMaxSamples;
  const int k = NS > 0 ? NS : k_rt;
  const bool floyd = explicit_idx == nullptr && bags.replace == 0;
  const bool paired = n <= kPairedMaxReads;
  // no bag of k distinct reads exists for n < k (np.random.choice raises in the reference, whose datasets drop such
  // sites, utils/data_utils.py:129); an empty site has no bag at all
  const bool no_bag = n == 0u || (floyd && n < static_cast<uint32_t>(k));
  float v = 0.0f;
  for (int r = 0; r < rounds; ++r) {
    uint32_t pick[kCap];
    float y;
    if (no_bag) {
      y = __int_as_float(0x7fc00000);
    } else {
      if (explicit_idx != nullptr) {
#pragma unroll
        for (int s = 0; s < k; ++s) pick[s] = min(static_cast<uint32_t>(explicit_idx[r * explicit_round_stride + s]), n - 1u);
      } else if (floyd) {
        floyd_bag<NS>(g, n, k, pick);
      } else {
        uint32_t pending = 0;
#pragma unroll
        for (int s = 0; s < k; ++s) {
          if (!paired) pick[s] = __umulhi(g.next(), n);
          else if ((s & 1) == 0) g.next_pair(n, pick[s], pending);
          else pick[s] = pending;
        }
      }
      if (bags.pool == kPoolProd) {
        float prod = 1.0f;
#pragma unroll
        for (int s = 0; s < k; ++s) prod *= 1.0f - pbase[pick[s]];
        y = 1.0f - prod;
      } else if (bags.pool == kPoolMean) {
        float sum = 0.0f;
#pragma unroll
        for (int s = 0; s < k; ++s) sum += pbase[pick[s]];
        y = sum / static_cast<float>(k);
      } else {
        float m = pbase[pick[0]];
#pragma unroll
        for (int s = 1; s < k; ++s) m = fmaxf(m, pbase[pick[s]]);
        y = m;
      }
    }
    if (bag_out != nullptr) bag_out[32 * r] = y;
    v += y;
  }
  return v;
}

// ---- feature staging geometry: rows [r0 + chunk*kChunkReads, ...) of feats as 16-byte granules -----------
struct Span {
  unsigned long long b0, b1;   // exact byte range of the rows
  unsigned long long g0, g1;   // granule range moved by the bulk copy (g1 clamped to the buffer's last full granule)
};
__device__ __forceinline__ Span chunk_span(const KernelArgs& a, long long r0, int nr, int chunk) {
  Span sp;
  const long long ra = r0 + static_cast<long long>(chunk) * kChunkReads;
  const int rows = min(kChunkReads, nr - chunk * kChunkReads);
  sp.b0 = static_cast<unsigned long long>(ra) * (kNSig * 4);
  sp.b1 = sp.b0 + static_cast<unsigned long long>(rows) * (kNSig * 4);
  sp.g0 = sp.b0 & ~15ull;
  sp.g1 = (sp.b1 + 15ull) & ~15ull;
  const unsigned long long gend = a.feats_bytes & ~15ull;
  if (sp.g1 > gend) sp.g1 = gend;
  return sp;
}
// ---- read encoder for R reads per thread (R = 2 full chunk rows, R = 1 for a warp whose second slot is empty) --
template <int R>
__device__ __forceinline__ void encode_reads(Smem& sm, const WeightImage& W, const KernelArgs& a,
                                             const float (&x)[kReadsPerThread][kNSig],
                                             const float* const (&cs)[kReadsPerThread], const int (&site_l)[kReadsPerThread],
                                             const bool (&valid)[kReadsPerThread], long long r0, int cbase, int tid,
                                             int n_pairs, bool q_in_smem) {
  float2 acc[R][kH2 / 2];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < kH2 / 2; ++k) acc[r][k] = make_float2(W.b2[2 * k], W.b2[2 * k + 1]);

#pragma unroll kPairUnroll
  for (int p = 0; p < n_pairs; ++p) {
    const float4* wp = reinterpret_cast<const float4*>(W.pair[p]);
    const float4 u0 = wp[0], u1 = wp[1], u2 = wp[2], u3 = wp[3], u4 = wp[4];
    float2 h[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 t = *reinterpret_cast<const float2*>(cs[r] + 2 * p);
      t = ffma2(make_float2(u0.x, u0.y), x[r][0], t);
      t = ffma2(make_float2(u0.z, u0.w), x[r][1], t);
      t = ffma2(make_float2(u1.x, u1.y), x[r][2], t);
      t = ffma2(make_float2(u1.z, u1.w), x[r][3], t);
      t = ffma2(make_float2(u2.x, u2.y), x[r][4], t);
      t = ffma2(make_float2(u2.z, u2.w), x[r][5], t);
      t = ffma2(make_float2(u3.x, u3.y), x[r][6], t);
      t = ffma2(make_float2(u3.z, u3.w), x[r][7], t);
      t = ffma2(make_float2(u4.x, u4.y), x[r][8], t);
      h[r] = make_float2(fmaxf(t.x, 0.0f), fmaxf(t.y, 0.0f));
    }
#pragma unroll
    for (int k4 = 0; k4 < kH2 / 4; ++k4) {
      const float4 w = wp[kW2Off0 / 4 + k4];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r][2 * k4 + 0] = ffma2(make_float2(w.x, w.y), h[r].x, acc[r][2 * k4 + 0]);
        acc[r][2 * k4 + 1] = ffma2(make_float2(w.z, w.w), h[r].x, acc[r][2 * k4 + 1]);
      }
    }
#pragma unroll
    for (int k4 = 0; k4 < kH2 / 4; ++k4) {
      const float4 w = wp[kW2Off1 / 4 + k4];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r][2 * k4 + 0] = ffma2(make_float2(w.x, w.y), h[r].y, acc[r][2 * k4 + 0]);
        acc[r][2 * k4 + 1] = ffma2(make_float2(w.z, w.w), h[r].y, acc[r][2 * k4 + 1]);
      }
    }
  }

#pragma unroll
  for (int r = 0; r < R; ++r) {
    float z = W.b3;
#pragma unroll
    for (int k = 0; k < kH2 / 2; ++k) {
      z = fmaf(W.w3[2 * k], fmaxf(acc[r][k].x, 0.0f), z);
      z = fmaf(W.w3[2 * k + 1], fmaxf(acc[r][k].y, 0.0f), z);
    }
    const float p = 1.0f / (1.0f + expf(-z));
    if (valid[r]) {
      const int lr = cbase + r * kThreads + tid;
      a.read_prob[r0 + lr] = p;
      if (q_in_smem) sm.q[lr] = 1.0f - p;
      if (p >= a.read_threshold) atomicAdd(&sm.cnt[site_l[r]], 1);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// BAGS = false: inference (phase B = Monte-Carlo noisy-OR, draws with replacement).  BAGS = true: the validate()-style
// literal MIL forward (phase B = bag_rounds); `bags` is read by that instantiation only.
template <int NS, bool BAGS>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
#if M6A_WEIGHTS_CONST
mil_infer_kernel(const KernelArgs a, const __grid_constant__ WeightImage wparam, const BagArgs bags) {
#else
mil_infer_kernel(const KernelArgs a, const BagArgs bags) {
#endif
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
#if M6A_WEIGHTS_CONST
  // The weights are a kernel parameter (constant bank 0): ptxas feeds them to FFMA2 through uniform registers
  // (LDCU.64 UR, c[0x0][UR+imm]; FFMA2 R, R.F32, UR.F32x2, R), so they use neither shared-memory bandwidth nor
  // vector registers.
  const WeightImage& W = wparam;
#else
  const WeightImage& W = sm.w;
#endif
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  // ---- one-time: weights image -> shared memory by one bulk copy -------------------------------
  if (tid == 0) {
    mbar_init(&sm.bar_w, 1);
    mbar_init(&sm.bar_f, 1);
    fence_barrier_init();
  }
  __syncthreads();
#if M6A_WEIGHTS_CONST
  bool weights_ready = true;
#else
  if (tid == 0) {
    mbar_expect_tx(&sm.bar_w, static_cast<uint32_t>(sizeof(WeightImage)));
    bulk_g2s(&sm.w, a.model.image, static_cast<uint32_t>(sizeof(WeightImage)), &sm.bar_w);
  }
  bool weights_ready = false;
#endif
  uint32_t f_parity = 0;

  const int n_pairs = (a.model.h1 + 1) >> 1;
  const int n_blocks = a.n_blocks, ipl = a.iters_per_lane;
  const bool tma_ok = a.feats_tma_ok;
  const float n_iters_f = static_cast<float>(a.n_iters);

#if M6A_DYNAMIC_TILES
  __shared__ long long s_next_tile;
  unsigned long long* tile_counter = reinterpret_cast<unsigned long long*>(const_cast<long long*>(a.tile_bounds)) + a.n_tiles + 1;
  for (;;) {
   if (tid == 0) s_next_tile = static_cast<long long>(atomicAdd(tile_counter, 1ull));
   __syncthreads();
   const long long tile = s_next_tile;
   __syncthreads();   // every thread has its copy before thread 0 draws the next tile
   if (tile >= a.n_tiles) break;
#else
  for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
#endif
   // a tile normally holds <= kSitesPerTileMax sites; more (very short sites) are taken in slices
   const long long tile_s0 = a.tile_bounds[tile], tile_s1 = a.tile_bounds[tile + 1];
   for (long long s0 = tile_s0; s0 < tile_s1; s0 += kSitesPerTileMax) {
    const int ns = static_cast<int>(min(static_cast<long long>(kSitesPerTileMax), tile_s1 - s0));
    const long long r0 = a.read_off[s0];

    // ---- tile header: local CSR offsets, k-mer ids, counters ------------------------------------
    if (tid <= ns) sm.roff[tid] = static_cast<int>(a.read_off[s0 + tid] - r0);
    if (tid < ns) {
      sm.cnt[tid] = 0;
#pragma unroll
      for (int t = 0; t < kKmerPos; ++t) {
        int k = a.kmer_idx != nullptr ? a.kmer_idx[(s0 + tid) * kKmerPos + t] : 0;
        k = (a.model.n_kmer == 1) ? 0 : min(max(k, 0), a.model.n_kmer - 1);
        sm.kid[tid][t] = k;
      }
    }
    __syncthreads();
    const int nr = sm.roff[ns];                 // reads in this tile
    const bool q_in_smem = !BAGS && nr <= kQCap;   // bags pool the per-read p itself (read back from read_prob)
    const int n_chunks = (nr + kChunkReads - 1) / kChunkReads;

    // ---- feature staging: rows [ra, ra+rows) of feats -> sm.feat[head + i*9 + k] ---------------------
    auto stage_chunk = [&](int chunk) {
      const Span sp = chunk_span(a, r0, nr, chunk);
      const unsigned long long b0 = sp.b0, b1 = sp.b1, g0 = sp.g0, g1 = sp.g1;
      if (tma_ok) {
        if (tid == 0 && g1 > g0) {
          mbar_expect_tx(&sm.bar_f, static_cast<uint32_t>(g1 - g0));
          bulk_g2s(sm.feat, reinterpret_cast<const unsigned char*>(a.feats) + g0, static_cast<uint32_t>(g1 - g0),
                   &sm.bar_f);
        }
        if (b1 > g1) {  // bytes past the last full 16-byte granule of the buffer (very last chunk only)
          const unsigned long long src0 = max(g1, b0);
          const int nf = static_cast<int>((b1 - src0) >> 2);
          if (tid < nf) sm.feat[((src0 - g0) >> 2) + tid] = a.feats[(src0 >> 2) + tid];
        }
      } else {
        const int nf = static_cast<int>((b1 - b0) >> 2);
        const float* src = a.feats + (b0 >> 2);
        for (int i = tid; i < nf; i += kThreads) sm.feat[i] = src[i];
      }
    };

    if (n_chunks > 0) stage_chunk(0);

    // c_site[s][j] = ctab0[k0][j] + ctab1[k1][j] + ctab2[k2][j]   (coalesced over j, L2 resident).
    // Four elements per thread per batch, all 12 loads issued before the first add: one L2 round trip per batch.
    {
      const float* ctab = a.model.ctab;
      const size_t tstride = static_cast<size_t>(a.model.n_kmer) * kH1Max;
      const int total = ns * kH1Max;
      for (int base = 0; base < total; base += 4 * kThreads) {
        float v[4][3];
        int at[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = base + u * kThreads + tid;
          const bool ok = i < total;
          const int s = ok ? i / kH1Max : 0, j = ok ? i - s * kH1Max : 0;
          at[u] = ok ? s * kCStride + j : -1;
#pragma unroll
          for (int t = 0; t < kKmerPos; ++t)
            v[u][t] = ok ? __ldg(ctab + t * tstride + static_cast<size_t>(sm.kid[s][t]) * kH1Max + j) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (at[u] >= 0) (&sm.csite[0][0])[at[u]] = (v[u][0] + v[u][1]) + v[u][2];
      }
    }
    if (!weights_ready) {
      mbar_wait(&sm.bar_w, 0);
      weights_ready = true;
    }

    // ---- phase A: read encoder ---------------------------------------------------------------------
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      const Span csp = chunk_span(a, r0, nr, chunk);
      const unsigned long long b0 = csp.b0, g0 = csp.g0, g1 = csp.g1;
      if (tma_ok && g1 > g0) {
        mbar_wait(&sm.bar_f, f_parity);
        f_parity ^= 1u;
      }
      __syncthreads();  // plain-load staging + csite visible
      const int head = tma_ok ? static_cast<int>((b0 - g0) >> 2) : 0;
      const int cbase = chunk * kChunkReads;

      float x[kReadsPerThread][kNSig];
      const float* cs[kReadsPerThread];
      int site_l[kReadsPerThread];
      bool valid[kReadsPerThread];
#pragma unroll
      for (int r = 0; r < kReadsPerThread; ++r) {
        const int lr = cbase + r * kThreads + tid;   // tile-local read index
        valid[r] = lr < nr;
        const int lrc = valid[r] ? lr : (nr - 1);
        int lo = 0, hi = ns;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (sm.roff[mid] <= lrc) lo = mid; else hi = mid;
        }
        site_l[r] = lo;
        cs[r] = sm.csite[lo];
        const float* xr = sm.feat + head + (lrc - cbase) * kNSig;
#pragma unroll
        for (int k = 0; k < kNSig; ++k) x[r][k] = xr[k];
      }
      __syncthreads();  // feature buffer free again
      if (chunk + 1 < n_chunks) stage_chunk(chunk + 1);   // overlaps the MLP below

      // Warp-uniform count of live read slots: ragged tiles leave the second slot (or both) of trailing warps empty,
      // and those warps run the one-read instantiation (or nothing) instead of computing on clamped rows.
      const bool any1 = __any_sync(0xffffffffu, valid[kReadsPerThread - 1]);
      const bool any0 = __any_sync(0xffffffffu, valid[0]);
      if (kReadsPerThread == 2 && any1) {
        encode_reads<kReadsPerThread>(sm, W, a, x, cs, site_l, valid, r0, cbase, tid, n_pairs, q_in_smem);
      } else if (any0) {
        encode_reads<1>(sm, W, a, x, cs, site_l, valid, r0, cbase, tid, n_pairs, q_in_smem);
      }
    }
    __syncthreads();  // q, cnt and (fallback) read_prob of the whole tile are visible

    // ---- phase B (bags): one bag per (site, pass), pooled; warp per (site, block of passes) --------------
    if constexpr (BAGS) {
      const int items = ns * n_blocks;
      for (int item = warp; item < items; item += kWarps) {
        const int sl_ = item / n_blocks, blk_ = item - sl_ * n_blocks;
        const int n = sm.roff[sl_ + 1] - sm.roff[sl_];
        float v;
        {
          // passes of this lane in block blk_: it = (blk_*ipl + r)*32 + lane, r < ipl, it < n_iters
          const long long it0 = static_cast<long long>(blk_) * ipl * 32 + lane;
          const long long left = (static_cast<long long>(a.n_iters) - it0 + 31) / 32;
          const int rounds = static_cast<int>(left < 0 ? 0 : (left > ipl ? ipl : left));
          Mwc64x g;
          g.seed(static_cast<uint32_t>(lane), static_cast<uint32_t>(blk_),
                 static_cast<unsigned long long>(a.site_id_base + s0 + sl_), a.seed);
          const size_t row = static_cast<size_t>(s0 + sl_) * a.n_iters + it0;
          const uint16_t* ex = a.sample_idx != nullptr ? a.sample_idx + row * a.n_samples : nullptr;
          float* bag_out = bags.bag_prob != nullptr ? bags.bag_prob + row : nullptr;
          v = bag_rounds<NS>(a.read_prob + r0 + sm.roff[sl_], static_cast<uint32_t>(n), g, rounds, a.n_samples, bags, ex,
                             static_cast<size_t>(32) * a.n_samples, bag_out);
        }
        v = warp_butterfly_sum(v);
        if (lane == 0) sm.partial[sl_][blk_] = v;
      }
    } else {
    // ---- phase B: Monte-Carlo noisy-OR ----------------------------------------------------------
      // items (site, block) are dealt round-robin to the warps, two at a time; (sl, blk) advance without a division
      const int items = ns * n_blocks;
      auto advance = [&](int& sl_, int& blk_) {
        blk_ += kWarps;
        while (blk_ >= n_blocks) { blk_ -= n_blocks; ++sl_; }
      };
      // rounds of this lane in block blk: it = (blk*ipl + k)*32 + lane, k < ipl, it < n_iters
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
          if (NS > 0 && q_in_smem && a.sample_idx == nullptr) {
            v = mc_lane_smem<NS>(sm.q + sm.roff[sl_], static_cast<uint32_t>(n), g, rounds);
          } else {
            const float* qbase = q_in_smem ? (sm.q + sm.roff[sl_]) : (a.read_prob + r0 + sm.roff[sl_]);
            const uint16_t* ex = a.sample_idx != nullptr
                                     ? a.sample_idx + (static_cast<size_t>(s0 + sl_) * a.n_iters + it0) * a.n_samples
                                     : nullptr;
            v = mc_lane_generic(qbase, !q_in_smem, static_cast<uint32_t>(n), g, rounds, a.n_samples, ex,
                                static_cast<size_t>(32) * a.n_samples);
          }
        }
        v = warp_butterfly_sum(v);
        if (lane == 0) sm.partial[sl_][blk_] = v;
      };

      int sl = 0, blk = warp;
      while (blk >= n_blocks) { blk -= n_blocks; ++sl; }
      for (int item = warp; item < items; item += 2 * kWarps) {
        int sl2 = sl, blk2 = blk;
        advance(sl2, blk2);
        const bool have2 = item + kWarps < items;
        const int na = sm.roff[sl + 1] - sm.roff[sl];
        const int nb = have2 ? sm.roff[sl2 + 1] - sm.roff[sl2] : 0;
        const bool fast2 = NS > 0 && q_in_smem && a.sample_idx == nullptr && have2 && na > 0 && nb > 0 &&
                           ((na <= static_cast<int>(kPairedMaxReads)) == (nb <= static_cast<int>(kPairedMaxReads)));
        if (fast2) {      // warp-uniform: two independent chains per lane, interleaved
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

    // ---- finalize -------------------------------------------------------------------------------
    if (tid < ns) {
      const int n = sm.roff[tid + 1] - sm.roff[tid];
      float s = 0.0f;
      for (int k = 0; k < n_blocks; ++k) s += sm.partial[tid][k];
      const size_t o = static_cast<size_t>(s0 + tid) * a.site_stride;
      a.site_prob[o] = n > 0 ? s / n_iters_f : __int_as_float(0x7fc00000);
      a.mod_count[o] = sm.cnt[tid];
    }
    __syncthreads();  // roff/cnt/partial are rewritten by the next tile
   }
  }

  if (!weights_ready) mbar_wait(&sm.bar_w, 0);  // never leave with a bulk copy in flight
}

// Prepass: tile t starts at the first site whose first feature row is >= t * tile_reads (binary search over read_off).
__global__ void tile_bounds_kernel(const int64_t* __restrict__ read_off, long long n_sites, long long n_tiles, int tile_reads,
                                   long long* __restrict__ tile_bounds) {
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t > n_tiles) return;
  if (t == n_tiles) {
    tile_bounds[t] = n_sites;
#if M6A_DYNAMIC_TILES
    tile_bounds[t + 1] = 0;                  // the tile counter of mil_infer_kernel
#endif
    return;
  }
  const long long target = t * tile_reads;
  long long lo = 0, hi = n_sites;            // first s in [0, n_sites) with read_off[s] >= target
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (read_off[mid] < target) lo = mid + 1; else hi = mid;
  }
  tile_bounds[t] = lo;
}

__global__ void sample_indices_kernel(uint64_t seed, uint64_t site_id, uint32_t n_reads, int n_iters, int n_samples,
                                      int n_blocks, int ipl, bool without_replacement, int32_t* __restrict__ out) {
  // one thread per (block, lane) stream, exactly the order phase B consumes it
  const int stream = blockIdx.x * blockDim.x + threadIdx.x;
  if (stream >= n_blocks * 32) return;
  const int blk = stream >> 5, lane = stream & 31;
  Mwc64x g;
  g.seed(static_cast<uint32_t>(lane), static_cast<uint32_t>(blk), site_id, seed);
  for (int k = 0; k < ipl; ++k) {
    const long long it = (static_cast<long long>(blk) * ipl + k) * 32 + lane;
    if (it >= n_iters) break;
    if (without_replacement) {   // Floyd bag, exactly as bag_rounds draws it
      uint32_t pick[kMaxSamples];
      floyd_bag<0>(g, n_reads, n_samples, pick);
      for (int s = 0; s < n_samples; ++s) out[it * n_samples + s] = static_cast<int32_t>(pick[s]);
      continue;
    }
    uint32_t pending = 0;
    for (int s = 0; s < n_samples; ++s) {
      uint32_t i;
      if (n_reads > kPairedMaxReads) i = __umulhi(g.next(), n_reads);
      else if ((s & 1) == 0) g.next_pair(n_reads, i, pending);
      else i = pending;
      out[it * n_samples + s] = static_cast<int32_t>(i);
    }
  }
}

// ---- host-side launchers ---------------------------------------------------------------------------
// occupancy / max-dynamic-smem attribute are per (device, kernel instantiation)
constexpr int kMaxDevices = 64;
constexpr int kVariants = 4;   // {n_samples == 20, generic} x {inference, bags}
static int g_max_ctas_per_sm[kMaxDevices][kVariants];
static bool g_attr_done[kMaxDevices][kVariants];

template <int NS, bool BAGS>
static cudaError_t launch_variant(const KernelArgs& a, const WeightImage* host_image, const BagArgs& bags, int n_sms,
                                  int dev, cudaStream_t stream, LaunchInfo* info) {
  auto kern = mil_infer_kernel<NS, BAGS>;
  constexpr int variant = (NS == 20 ? 0 : 1) + (BAGS ? 2 : 0);
  const int smem = static_cast<int>(sizeof(Smem));
  int& occ = g_max_ctas_per_sm[dev][variant];
  if (!g_attr_done[dev][variant]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem);
    if (e != cudaSuccess) return e;
    occ = nb > 0 ? nb : 1;
    g_attr_done[dev][variant] = true;
  }
  long long grid = static_cast<long long>(n_sms) * occ;
  if (grid > a.n_tiles) grid = a.n_tiles;
  if (grid < 1) grid = 1;
  if (info) {
    info->grid = static_cast<int>(grid);
    info->block = kThreads;
    info->smem_bytes = smem;
    info->tile_reads = a.tile_reads;
  }
#if M6A_WEIGHTS_CONST
  kern<<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(a, *host_image, bags);
#else
  (void)host_image;
  kern<<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(a, bags);
#endif
  return cudaGetLastError();
}

cudaError_t launch_mil_infer(const KernelArgs& a, const WeightImage* host_image, const BagArgs* bags, int n_sms,
                             cudaStream_t stream, LaunchInfo* info) {
  int dev = 0;
  cudaError_t de = cudaGetDevice(&dev);
  if (de != cudaSuccess) return de;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  const bool fast = (a.n_samples == 20);
  if (bags != nullptr) {
    return fast ? launch_variant<20, true>(a, host_image, *bags, n_sms, dev, stream, info)
                : launch_variant<0, true>(a, host_image, *bags, n_sms, dev, stream, info);
  }
  const BagArgs none = {nullptr, 1, kPoolProd};
  return fast ? launch_variant<20, false>(a, host_image, none, n_sms, dev, stream, info)
              : launch_variant<0, false>(a, host_image, none, n_sms, dev, stream, info);
}

cudaError_t launch_tile_bounds(const int64_t* read_off, long long n_sites, long long n_tiles, int tile_reads,
                               long long* tile_bounds, cudaStream_t stream) {
  const long long n = n_tiles + 1;
  tile_bounds_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(read_off, n_sites, n_tiles, tile_reads,
                                                                               tile_bounds);
  return cudaGetLastError();
}

cudaError_t launch_sample_indices(uint64_t seed, uint64_t site_id, uint32_t n_reads, int n_iters, int n_samples,
                                  bool without_replacement, int32_t* out, cudaStream_t stream) {
  int ipl, n_blocks;
  block_layout(n_iters, &ipl, &n_blocks);
  const int streams = n_blocks * 32;
  sample_indices_kernel<<<(streams + 127) / 128, 128, 0, stream>>>(seed, site_id, n_reads, n_iters, n_samples,
                                                                  n_blocks, ipl, without_replacement, out);
  return cudaGetLastError();
}

}  // namespace m6a
