// Fused MIL-inference kernel for sm_100a: read encoder -> q = 1-p in shared memory ->
// Monte-Carlo noisy-OR mean -> site probability, mod_count, per-read probability.
//
// Replaces, per tile of consecutive sites, the reference's
//   model.get_read_representation + pooling_filter.probability_layer   utils/inference_utils.py:35-37
//   group_results / mod_ratio                                          utils/inference_utils.py:48-53
//   calculate_site_proba (_calculate_site_proba)                       utils/inference_utils.py:54,74-104
//
// Work decomposition (one persistent CTA loops over tiles; a tile = G consecutive sites):
//   stage   the tile's feature rows are one contiguous byte range of `feats` -> one cp.async.bulk
//           (TMA, mbarrier completion) per chunk of kChunkReads rows into shared memory
//   phase A thread-per-read (kReadsPerThread reads per thread): h1-step fused loop
//             h_j = relu(c_site[j] + w1[j,0:9].x);  acc[0:32] += w2[:,j] * h_j
//           weights are read from shared memory at warp-uniform addresses (LDS.128 broadcast);
//           the k-mer embedding part of Linear-1 is a per-site constant c_site (m6a_layout.h)
//   phase B warp-per-(site, slab of 32 iterations): every lane runs one MC iteration: 5 Philox4x32-10
//           calls -> 20 indices -> product of q[idx] from shared memory; butterfly reduce per slab
//   final   ordered sum of slab partials / n_iters -> site_prob
// The summation order is a function of n_iters only (not of tiling, grid or GPU count), so a site's
// result is bit-identical however the sites are sharded.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "m6a_kernel.h"
#include "m6a_philox.cuh"

namespace m6a {

struct Smem {
  WeightImage w;
  alignas(16) float feat[kChunkReads * kNSig + 8];  // + unaligned head (<=3 floats) + tail round-up
  float csite[kSitesPerTileMax][kCStride];
  float q[kQCap];
  float partial[kSitesPerTileMax][kSlabCap];
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

// ---- one Monte-Carlo iteration: 1 - prod_{s<n_samples} q[idx_s] ---------------------------------
template <int NS>
__device__ __forceinline__ float mc_iteration_philox(const float* __restrict__ qs, uint32_t n, uint32_t it,
                                                     uint32_t site_lo, uint32_t site_hi, uint32_t k0,
                                                     uint32_t k1, int n_samples_rt) {
  const int ns = NS > 0 ? NS : n_samples_rt;
  float prod = 1.0f;
  if (NS > 0) {
#pragma unroll
    for (int c = 0; c < (NS + 3) / 4; ++c) {
      const Philox4 r = philox4x32_10(static_cast<uint32_t>(c), it, site_lo, site_hi, k0, k1);
      if (4 * c + 0 < NS) prod *= qs[__umulhi(r.x, n)];
      if (4 * c + 1 < NS) prod *= qs[__umulhi(r.y, n)];
      if (4 * c + 2 < NS) prod *= qs[__umulhi(r.z, n)];
      if (4 * c + 3 < NS) prod *= qs[__umulhi(r.w, n)];
    }
  } else {
    for (int c = 0; 4 * c < ns; ++c) {
      const Philox4 r = philox4x32_10(static_cast<uint32_t>(c), it, site_lo, site_hi, k0, k1);
      prod *= qs[__umulhi(r.x, n)];
      if (4 * c + 1 < ns) prod *= qs[__umulhi(r.y, n)];
      if (4 * c + 2 < ns) prod *= qs[__umulhi(r.z, n)];
      if (4 * c + 3 < ns) prod *= qs[__umulhi(r.w, n)];
    }
  }
  return 1.0f - prod;
}

// q read back from global read_prob (tile too large for the shared q table)
template <bool kFromProb>
__device__ __forceinline__ float q_at(const float* base, uint32_t i) {
  return kFromProb ? 1.0f - base[i] : base[i];
}

__device__ __forceinline__ float mc_iteration_generic(const float* qbase, bool from_prob, uint32_t n, uint32_t it,
                                                      uint32_t site_lo, uint32_t site_hi, uint32_t k0, uint32_t k1,
                                                      int ns, const uint16_t* __restrict__ explicit_idx) {
  float prod = 1.0f;
  if (explicit_idx != nullptr) {
    for (int s = 0; s < ns; ++s) {
      const uint32_t i = explicit_idx[s];
      prod *= from_prob ? q_at<true>(qbase, i) : q_at<false>(qbase, i);
    }
  } else {
    for (int c = 0; 4 * c < ns; ++c) {
      const Philox4 r = philox4x32_10(static_cast<uint32_t>(c), it, site_lo, site_hi, k0, k1);
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (4 * c + u < ns) {
          const uint32_t i = __umulhi(w[u], n);
          prod *= from_prob ? q_at<true>(qbase, i) : q_at<false>(qbase, i);
        }
      }
    }
  }
  return 1.0f - prod;
}

__device__ __forceinline__ float warp_butterfly_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// -------------------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(kThreads, 2)
mil_infer_kernel(const KernelArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
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
  if (tid == 0) {
    mbar_expect_tx(&sm.bar_w, static_cast<uint32_t>(sizeof(WeightImage)));
    bulk_g2s(&sm.w, a.model.image, static_cast<uint32_t>(sizeof(WeightImage)), &sm.bar_w);
  }
  uint32_t f_parity = 0;
  bool weights_ready = false;

  const int h1 = a.model.h1;
  const uint32_t k0 = static_cast<uint32_t>(a.seed), k1 = static_cast<uint32_t>(a.seed >> 32);
  const int n_slabs = a.n_slabs, ipl = a.iters_per_lane;
  const bool tma_ok = a.feats_tma_ok;
  const float inv_iters_den = static_cast<float>(a.n_iters);

  for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const long long s0 = tile * a.sites_per_tile;
    const int ns = static_cast<int>(min(static_cast<long long>(a.sites_per_tile), a.n_sites - s0));
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
    const bool q_in_smem = nr <= kQCap;
    const int n_chunks = (nr + kChunkReads - 1) / kChunkReads;

    // ---- stage chunk 0 (TMA) while c_site is being computed --------------------------------------
    auto stage_chunk = [&](int chunk) {
      // rows [ra, rb) of feats -> sm.feat[head + i*9 + k]
      const long long ra = r0 + static_cast<long long>(chunk) * kChunkReads;
      const int rows = min(kChunkReads, nr - chunk * kChunkReads);
      const unsigned long long b0 = static_cast<unsigned long long>(ra) * (kNSig * 4);
      const unsigned long long b1 = b0 + static_cast<unsigned long long>(rows) * (kNSig * 4);
      if (tma_ok) {
        const unsigned long long g0 = b0 & ~15ull;
        unsigned long long g1 = (b1 + 15ull) & ~15ull;
        const unsigned long long gend = a.feats_bytes & ~15ull;
        if (g1 > gend) g1 = gend;
        if (tid == 0 && g1 > g0) {
          mbar_expect_tx(&sm.bar_f, static_cast<uint32_t>(g1 - g0));
          bulk_g2s(sm.feat, reinterpret_cast<const unsigned char*>(a.feats) + g0, static_cast<uint32_t>(g1 - g0),
                   &sm.bar_f);
        }
        // bytes past the last 16-byte granule of the buffer (only the very last chunk can have them)
        if (b1 > g1) {
          const int nf = static_cast<int>((b1 - max(g1, b0)) >> 2);
          const unsigned long long src0 = max(g1, b0);
          if (tid < nf) sm.feat[((src0 - g0) >> 2) + tid] = a.feats[(src0 >> 2) + tid];
        }
      } else {
        const int nf = rows * kNSig;
        const float* src = a.feats + ra * kNSig;
        for (int i = tid; i < nf; i += kThreads) sm.feat[i] = src[i];
      }
    };
    auto chunk_has_tma = [&](int chunk) -> bool {
      if (!tma_ok) return false;
      const long long ra = r0 + static_cast<long long>(chunk) * kChunkReads;
      const int rows = min(kChunkReads, nr - chunk * kChunkReads);
      const unsigned long long b0 = static_cast<unsigned long long>(ra) * (kNSig * 4);
      const unsigned long long b1 = b0 + static_cast<unsigned long long>(rows) * (kNSig * 4);
      unsigned long long g1 = (b1 + 15ull) & ~15ull;
      const unsigned long long gend = a.feats_bytes & ~15ull;
      if (g1 > gend) g1 = gend;
      return g1 > (b0 & ~15ull);
    };
    auto chunk_head = [&](int chunk) -> int {
      if (!tma_ok) return 0;
      const unsigned long long b0 =
          static_cast<unsigned long long>(r0 + static_cast<long long>(chunk) * kChunkReads) * (kNSig * 4);
      return static_cast<int>((b0 & 15ull) >> 2);
    };

    if (n_chunks > 0) stage_chunk(0);

    // c_site[s][j] = ctab0[k0][j] + ctab1[k1][j] + ctab2[k2][j]   (coalesced over j, L2 resident)
    {
      const float* ctab = a.model.ctab;
      const size_t tstride = static_cast<size_t>(a.model.n_kmer) * kH1Max;
      for (int i = tid; i < ns * kH1Max; i += kThreads) {
        const int s = i / kH1Max, j = i - s * kH1Max;
        const float c = __ldg(ctab + static_cast<size_t>(sm.kid[s][0]) * kH1Max + j) +
                        __ldg(ctab + tstride + static_cast<size_t>(sm.kid[s][1]) * kH1Max + j) +
                        __ldg(ctab + 2 * tstride + static_cast<size_t>(sm.kid[s][2]) * kH1Max + j);
        sm.csite[s][j] = c;
      }
    }
    if (!weights_ready) {
      mbar_wait(&sm.bar_w, 0);
      weights_ready = true;
    }

    // ---- phase A: read encoder ---------------------------------------------------------------------
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      if (chunk_has_tma(chunk)) {
        mbar_wait(&sm.bar_f, f_parity);
        f_parity ^= 1u;
      }
      __syncthreads();  // plain-load staging + csite visible
      const int head = chunk_head(chunk);
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

      float acc[kReadsPerThread][kH2];
#pragma unroll
      for (int r = 0; r < kReadsPerThread; ++r)
#pragma unroll
        for (int k = 0; k < kH2; ++k) acc[r][k] = sm.w.b2[k];

#pragma unroll 2
      for (int j = 0; j < h1; ++j) {
        const float4* wj = reinterpret_cast<const float4*>(sm.w.l12[j]);
        const float4 wa = wj[0], wb = wj[1], wc = wj[2];
        float h[kReadsPerThread];
#pragma unroll
        for (int r = 0; r < kReadsPerThread; ++r) {
          float t = cs[r][j];
          t = fmaf(wa.x, x[r][0], t);
          t = fmaf(wa.y, x[r][1], t);
          t = fmaf(wa.z, x[r][2], t);
          t = fmaf(wa.w, x[r][3], t);
          t = fmaf(wb.x, x[r][4], t);
          t = fmaf(wb.y, x[r][5], t);
          t = fmaf(wb.z, x[r][6], t);
          t = fmaf(wb.w, x[r][7], t);
          t = fmaf(wc.x, x[r][8], t);
          h[r] = fmaxf(t, 0.0f);
        }
#pragma unroll
        for (int k4 = 0; k4 < kH2 / 4; ++k4) {
          const float4 w = wj[kW2Off / 4 + k4];
#pragma unroll
          for (int r = 0; r < kReadsPerThread; ++r) {
            acc[r][4 * k4 + 0] = fmaf(w.x, h[r], acc[r][4 * k4 + 0]);
            acc[r][4 * k4 + 1] = fmaf(w.y, h[r], acc[r][4 * k4 + 1]);
            acc[r][4 * k4 + 2] = fmaf(w.z, h[r], acc[r][4 * k4 + 2]);
            acc[r][4 * k4 + 3] = fmaf(w.w, h[r], acc[r][4 * k4 + 3]);
          }
        }
      }

#pragma unroll
      for (int r = 0; r < kReadsPerThread; ++r) {
        float z = sm.w.b3;
#pragma unroll
        for (int k = 0; k < kH2; ++k) z = fmaf(sm.w.w3[k], fmaxf(acc[r][k], 0.0f), z);
        const float p = 1.0f / (1.0f + expf(-z));
        if (valid[r]) {
          const int lr = cbase + r * kThreads + tid;
          a.read_prob[r0 + lr] = p;
          if (q_in_smem) sm.q[lr] = 1.0f - p;
          if (p >= a.read_threshold) atomicAdd(&sm.cnt[site_l[r]], 1);
        }
      }
    }
    __syncthreads();  // q, cnt and (fallback) read_prob of the whole tile are visible

    // ---- phase B: Monte-Carlo noisy-OR ----------------------------------------------------------
    {
      const int items = ns * n_slabs;
      for (int item = warp; item < items; item += kWarps) {
        const int sl = item / n_slabs, slab = item - sl * n_slabs;
        const int n = sm.roff[sl + 1] - sm.roff[sl];
        float v = 0.0f;
        if (n > 0) {
          const unsigned long long gsite = static_cast<unsigned long long>(a.site_id_base + s0 + sl);
          const uint32_t site_lo = static_cast<uint32_t>(gsite), site_hi = static_cast<uint32_t>(gsite >> 32);
          for (int jj = 0; jj < ipl; ++jj) {
            const long long it = (static_cast<long long>(slab) * ipl + jj) * 32 + lane;
            if (it < a.n_iters) {
              if (NS > 0 && q_in_smem && a.sample_idx == nullptr) {
                v += mc_iteration_philox<NS>(sm.q + sm.roff[sl], static_cast<uint32_t>(n), static_cast<uint32_t>(it),
                                             site_lo, site_hi, k0, k1, a.n_samples);
              } else {
                const float* qbase = q_in_smem ? (sm.q + sm.roff[sl]) : (a.read_prob + r0 + sm.roff[sl]);
                const uint16_t* ex =
                    a.sample_idx != nullptr
                        ? a.sample_idx + (static_cast<size_t>(s0 + sl) * a.n_iters + it) * a.n_samples
                        : nullptr;
                v += mc_iteration_generic(qbase, !q_in_smem, static_cast<uint32_t>(n), static_cast<uint32_t>(it),
                                          site_lo, site_hi, k0, k1, a.n_samples, ex);
              }
            }
          }
        }
        v = warp_butterfly_sum(v);
        if (lane == 0) sm.partial[sl][slab] = v;
      }
    }
    __syncthreads();

    // ---- finalize -------------------------------------------------------------------------------
    if (tid < ns) {
      const int n = sm.roff[tid + 1] - sm.roff[tid];
      float s = 0.0f;
      for (int k = 0; k < n_slabs; ++k) s += sm.partial[tid][k];
      a.site_prob[s0 + tid] = n > 0 ? s / inv_iters_den : __int_as_float(0x7fc00000);
      a.mod_count[s0 + tid] = sm.cnt[tid];
    }
    __syncthreads();  // roff/cnt/partial are rewritten by the next tile
  }

  if (!weights_ready) mbar_wait(&sm.bar_w, 0);  // never leave with a bulk copy in flight
}

__global__ void philox_indices_kernel(uint64_t seed, uint64_t site_id, uint32_t n_reads, int n_iters, int n_samples,
                                      int32_t* __restrict__ out) {
  const int n_calls = (n_samples + 3) / 4;
  const long long total = static_cast<long long>(n_iters) * n_calls;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int it = static_cast<int>(i / n_calls), c = static_cast<int>(i - static_cast<long long>(it) * n_calls);
    const Philox4 r = philox4x32_10(static_cast<uint32_t>(c), static_cast<uint32_t>(it), static_cast<uint32_t>(site_id),
                                    static_cast<uint32_t>(site_id >> 32), static_cast<uint32_t>(seed),
                                    static_cast<uint32_t>(seed >> 32));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    for (int u = 0; u < 4; ++u)
      if (4 * c + u < n_samples) out[static_cast<long long>(it) * n_samples + 4 * c + u] = __umulhi(w[u], n_reads);
  }
}

// ---- host-side launchers ---------------------------------------------------------------------------
static int g_max_ctas_per_sm[2] = {-1, -1};

cudaError_t launch_mil_infer(const KernelArgs& a, int n_sms, cudaStream_t stream, LaunchInfo* info) {
  const bool fast = (a.n_samples == 20);
  auto kfast = mil_infer_kernel<20>;
  auto kgen = mil_infer_kernel<0>;
  const void* fn = fast ? reinterpret_cast<const void*>(kfast) : reinterpret_cast<const void*>(kgen);
  const int smem = static_cast<int>(sizeof(Smem));
  int& occ = g_max_ctas_per_sm[fast ? 0 : 1];
  if (occ < 0) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int nb = 0;
    e = fast ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kfast, kThreads, smem)
             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kgen, kThreads, smem);
    if (e != cudaSuccess) return e;
    occ = nb > 0 ? nb : 1;
  }
  long long grid = static_cast<long long>(n_sms) * occ;
  if (grid > a.n_tiles) grid = a.n_tiles;
  if (grid < 1) grid = 1;
  if (info) {
    info->grid = static_cast<int>(grid);
    info->block = kThreads;
    info->smem_bytes = smem;
    info->sites_per_tile = a.sites_per_tile;
  }
  if (fast)
    kfast<<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(a);
  else
    kgen<<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_philox_indices(uint64_t seed, uint64_t site_id, uint32_t n_reads, int n_iters, int n_samples,
                                  int32_t* out, cudaStream_t stream) {
  philox_indices_kernel<<<64, 256, 0, stream>>>(seed, site_id, n_reads, n_iters, n_samples, out);
  return cudaGetLastError();
}

}  // namespace m6a
