// Monte-Carlo noisy-OR building blocks of phase B (shared by the product kernel m6a_kernel.cu and the experimental fused
// tensor-core kernel): one lane's share of a (site, block) = its rounds of 1 - prod_{s < n_samples} q[idx_s] on the
// lane's MWC64X stream (m6a_rng.cuh).  Specification of the streams and of the summation order: oracle/philox.py.
#pragma once
#include <stdint.h>

#include "m6a_rng.cuh"

namespace m6a {

__device__ __forceinline__ uint32_t mc_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_butterfly_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// ---- one lane's share of a (site, block): ipl iterations of 1 - prod_{s<n_samples} q[idx_s] ------
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");   // ordered after the phase barrier
  return v;
}

// v + (1 - prod) and the products themselves as separately rounded IEEE operations (never contracted into an FMA): every
// pooling path -- one chain, two chains, packed products, the global-memory fallback -- evaluates the same float32
// expression as the oracle, so a site's result does not depend on which path its slab takes.
__device__ __forceinline__ float mc_accumulate(float v, float prod) { return __fadd_rn(v, __fsub_rn(1.0f, prod)); }

// One draw step of a lane's chain in the paired (n <= 256, two indices per word) or single regime.
template <bool PAIRED>
__device__ __forceinline__ void mc_step(uint32_t qaddr, uint32_t n, Mwc64x& g, float& prod, bool second_of_pair_wanted = true) {
  if (PAIRED) {
    uint32_t i1, i2;
    g.next_pair(n, i1, i2);
    prod = __fmul_rn(prod, lds_f32(qaddr + (i1 << 2)));      // IMAD.WIDE, IMAD.HI, 2 x (LEA, LDS, FMUL) per word
    if (second_of_pair_wanted) prod = __fmul_rn(prod, lds_f32(qaddr + (i2 << 2)));
  } else {
    prod = __fmul_rn(prod, lds_f32(qaddr + (__umulhi(g.next(), n) << 2)));   // IMAD.HI, LEA, LDS, FMUL
  }
}

// rounds [k0, k1) of ONE chain, accumulated into v in round order (the canonical summation order)
template <int NS, bool PAIRED>
__device__ __forceinline__ void mc_rounds(uint32_t qaddr, uint32_t n, Mwc64x& g, int k0, int k1, float& v) {
  constexpr int kSteps = PAIRED ? NS / 2 : NS;
  for (int k = k0; k < k1; ++k) {
    float prod = 1.0f;
#pragma unroll
    for (int s = 0; s < kSteps; ++s) mc_step<PAIRED>(qaddr, n, g, prod);
    if (PAIRED && (NS & 1)) mc_step<PAIRED>(qaddr, n, g, prod, false);
    v = mc_accumulate(v, prod);
  }
}

// rounds [0, kc) of TWO independent chains interleaved instruction by instruction: phase B is bound by the latency of
// the serial MWC chain (wide multiply -> carry adds -> next multiply), and it has registers to spare
template <int NS, bool PAIRED>
__device__ __forceinline__ void mc_rounds_x2(uint32_t qa, uint32_t na, Mwc64x& ga, float& va, uint32_t qb, uint32_t nb,
                                             Mwc64x& gb, float& vb, int kc) {
  constexpr int kSteps = PAIRED ? NS / 2 : NS;
  for (int k = 0; k < kc; ++k) {
    float pa = 1.0f, pb = 1.0f;
#pragma unroll
    for (int s = 0; s < kSteps; ++s) {
      mc_step<PAIRED>(qa, na, ga, pa);
      mc_step<PAIRED>(qb, nb, gb, pb);
    }
    if (PAIRED && (NS & 1)) {
      mc_step<PAIRED>(qa, na, ga, pa, false);
      mc_step<PAIRED>(qb, nb, gb, pb, false);
    }
    va = mc_accumulate(va, pa);
    vb = mc_accumulate(vb, pb);
  }
}

template <int NS>
__device__ __forceinline__ float mc_lane_smem(const float* __restrict__ qs, uint32_t n, Mwc64x& g, int rounds) {
  const uint32_t qaddr = mc_smem_u32(qs);     // shared-space byte address of the site's q table
  float v = 0.0f;
  if (n <= kPairedMaxReads) mc_rounds<NS, true>(qaddr, n, g, 0, rounds, v);   // warp-uniform branch
  else mc_rounds<NS, false>(qaddr, n, g, 0, rounds, v);
  return v;
}

// two (site, block) items of the same index regime at once; identical results to two mc_lane_smem calls
template <int NS>
__device__ __forceinline__ void mc_lane_smem_x2(const float* qsa, uint32_t na, Mwc64x& ga, int ra, float& va,
                                                const float* qsb, uint32_t nb, Mwc64x& gb, int rb, float& vb) {
  const uint32_t qa = mc_smem_u32(qsa), qb = mc_smem_u32(qsb);
  const int kc = min(ra, rb);
  va = 0.0f;
  vb = 0.0f;
  if (na <= kPairedMaxReads) {
    mc_rounds_x2<NS, true>(qa, na, ga, va, qb, nb, gb, vb, kc);
    mc_rounds<NS, true>(qa, na, ga, kc, ra, va);
    mc_rounds<NS, true>(qb, nb, gb, kc, rb, vb);
  } else {
    mc_rounds_x2<NS, false>(qa, na, ga, va, qb, nb, gb, vb, kc);
    mc_rounds<NS, false>(qa, na, ga, kc, ra, va);
    mc_rounds<NS, false>(qb, nb, gb, kc, rb, vb);
  }
}

// NC (site, block) items of the same index regime at once, chains interleaved instruction by instruction (the tensor-core
// kernel has fewer Monte-Carlo warps than the CUDA-core kernel and hides the serial MWC latency inside each warp instead).
// Identical results to NC mc_lane_smem calls: every chain accumulates its own rounds in round order.
#ifndef M6A_MC_PACKED_MUL
#define M6A_MC_PACKED_MUL 1
#endif
template <int NS, bool PAIRED, int NC>
__device__ __forceinline__ void mc_rounds_xn(const uint32_t (&qa)[NC], const uint32_t (&n)[NC], Mwc64x (&g)[NC],
                                             const int (&rounds)[NC], float (&v)[NC]) {
  constexpr int kSteps = PAIRED ? NS / 2 : NS;
  int kc = rounds[0];
#pragma unroll
  for (int i = 1; i < NC; ++i) kc = min(kc, rounds[i]);
  if constexpr (M6A_MC_PACKED_MUL && NC == 2 && PAIRED && (NS & 1) == 0) {
    // the products of the two chains as one packed multiply (FMUL2: two IEEE multiplies, one issue slot)
    for (int k = 0; k < kc; ++k) {
      float2 p;
#pragma unroll
      for (int s = 0; s < kSteps; ++s) {
        uint32_t a1, a2, b1, b2;
        g[0].next_pair(n[0], a1, a2);
        g[1].next_pair(n[1], b1, b2);
        const float2 q1 = make_float2(lds_f32(qa[0] + (a1 << 2)), lds_f32(qa[1] + (b1 << 2)));
        p = s == 0 ? q1 : __fmul2_rn(p, q1);                       // (1.0f * q is q: the first factor starts the product)
        p = __fmul2_rn(p, make_float2(lds_f32(qa[0] + (a2 << 2)), lds_f32(qa[1] + (b2 << 2))));
      }
      v[0] = mc_accumulate(v[0], p.x);
      v[1] = mc_accumulate(v[1], p.y);
    }
  } else {
    for (int k = 0; k < kc; ++k) {
      float p[NC];
#pragma unroll
      for (int i = 0; i < NC; ++i) p[i] = 1.0f;
#pragma unroll
      for (int s = 0; s < kSteps; ++s) {
#pragma unroll
        for (int i = 0; i < NC; ++i) mc_step<PAIRED>(qa[i], n[i], g[i], p[i]);
      }
      if (PAIRED && (NS & 1)) {
#pragma unroll
        for (int i = 0; i < NC; ++i) mc_step<PAIRED>(qa[i], n[i], g[i], p[i], false);
      }
#pragma unroll
      for (int i = 0; i < NC; ++i) v[i] = mc_accumulate(v[i], p[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) mc_rounds<NS, PAIRED>(qa[i], n[i], g[i], kc, rounds[i], v[i]);   // ragged tails
}

// generic path: any n_samples, q from shared memory or 1 - read_prob from global, optional explicit indices
__device__ __forceinline__ float mc_lane_generic(const float* qbase, bool from_prob, uint32_t n, Mwc64x& g, int rounds,
                                                 int ns, const uint16_t* __restrict__ explicit_idx,
                                                 size_t explicit_round_stride) {
  float v = 0.0f;
  const bool paired = n <= kPairedMaxReads;
  for (int k = 0; k < rounds; ++k) {
    float prod = 1.0f;
    uint32_t pending = 0;
    for (int s = 0; s < ns; ++s) {
      uint32_t i;
      if (explicit_idx != nullptr) {
        i = explicit_idx[k * explicit_round_stride + s];
      } else if (!paired) {
        i = __umulhi(g.next(), n);
      } else if ((s & 1) == 0) {
        g.next_pair(n, i, pending);
      } else {
        i = pending;
      }
      prod = __fmul_rn(prod, from_prob ? __fsub_rn(1.0f, qbase[i]) : qbase[i]);
    }
    v = mc_accumulate(v, prod);
  }
  return v;
}

}  // namespace m6a
