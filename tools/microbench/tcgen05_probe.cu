// Bring-up probe for the tcgen05 building blocks of the EXPERIMENTAL tensor-core encoder
// (m6anet_b200/csrc/m6a_tc.cuh).  One CTA of 128 threads, every step checked against the host:
//   1. TMEM alloc / tcgen05.st / tcgen05.ld round trip (lane quadrants, column addressing)
//   2. one tcgen05.mma kind::tf32 SS (A, B from shared memory, K-major SWIZZLE_NONE descriptors), M128 x N x K8
//   3. the same product with A from TMEM (TS)
//   4. accumulation over two K-steps (descriptor advance by kStep), N = 160 and N = 32
//   5. what the tensor core does with the low 13 mantissa bits of a 32-bit operand (truncate / round / use them)
// Inputs are small integers (exact in TF32), so every expected value is exact.  Every mbarrier wait is bounded.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I m6anet_b200/csrc \
//        -o tools/microbench/tcgen05_probe tools/microbench/tcgen05_probe.cu && timeout 60 tools/microbench/tcgen05_probe
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "m6a_tc.cuh"

using namespace m6a::tc;
#define tc_mbar_init mbar_init
#define tc_fence_barrier_init fence_barrier_init
#define tc_fence_proxy_async fence_proxy_async
#define tc_fence_before fence_before
#define tc_fence_after fence_after
#define tc_wait_st wait_st
#define tc_wait_ld wait_ld
#define tc_commit mma_commit
#define tc_smem_u32 smem_u32
#define tc_mbar_wait(b, p) mbar_wait(b, p, 99)

constexpr int kRows = 128;
constexpr int kNMax = 160;
constexpr int kKTot = 16;   // two K-steps

struct ProbeSmem {
  alignas(128) float a[kKTot / 4][kRows][4];
  alignas(128) float b[kKTot / 4][kNMax][4];
  uint32_t tmem_base;
  alignas(8) unsigned long long bar;
};

// mode 0: st/ld round trip; 1: SS one K-step; 2: TS one K-step; 3: SS two K-steps
__global__ void __launch_bounds__(128, 1) probe_kernel(int mode, int n, const float* __restrict__ A, const float* __restrict__ B,
                                                      float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char raw[];
  ProbeSmem& sm = *reinterpret_cast<ProbeSmem*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  // operands -> shared memory in the K-major no-swizzle layout: element (row, k) at [k / 4][row][k % 4]
  for (int i = tid; i < kRows * kKTot; i += 128) sm.a[(i % kKTot) / 4][i / kKTot][i % 4] = A[i];
  for (int i = tid; i < n * kKTot; i += 128) {
    const int row = i / kKTot, k = i % kKTot;
    (&sm.b[0][0][0])[(k / 4) * n * 4 + row * 4 + (k % 4)] = B[i];      // rows = n for this launch (LBO = n * 16)
  }
  if (tid == 0) {
    tc_mbar_init(&sm.bar, 1);
    tc_fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 256);
  tc_fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  const uint32_t d_acc = tmem, a_tm = tmem + 192;     // accumulators [0,160), A-in-TMEM staging [192, 208)
  const uint32_t idesc = make_idesc(128, n);

  if (mode == 0) {        // write lane*1000 + column, read it back
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(static_cast<float>(tid * 1000 + i));
    tmem_st32(tmem + lane_base + 32, v);
    tc_wait_st();
    uint32_t w[32];
    tmem_ld32(tmem + lane_base + 32, w);
    tc_wait_ld();
    for (int i = 0; i < 32; ++i) D[tid * 32 + i] = __uint_as_float(w[i]);
  } else {
    if (mode == 2) {      // A (first K-step) -> TMEM columns [192, 200): lane = row, column = k
      uint32_t v[16];
      for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(i < 8 ? A[tid * kKTot + i] : 0.0f);
      tmem_st16(a_tm + lane_base, v);
      tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t sa = tc_smem_u32(sm.a), sb = tc_smem_u32(sm.b);
      const uint32_t lbo_b = n * 16, step_b = 2 * lbo_b;
      const uint64_t da = make_desc(sa, kRows * 16, 128), db = make_desc(sb, lbo_b, 128);
      if (mode == 1) {
        mma_ss(d_acc, da, db, idesc, 0u);
      } else if (mode == 2) {
        mma_ts(d_acc, a_tm, db, idesc, 0u);
      } else {
        mma_ss(d_acc, da, db, idesc, 0u);
        mma_ss(d_acc, make_desc(sa + 2 * kRows * 16, kRows * 16, 128), make_desc(sb + step_b, lbo_b, 128), idesc, 1u);
      }
      tc_commit(&sm.bar);
    }
    tc_mbar_wait(&sm.bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < n; c0 += 32) {
      uint32_t w[32];
      tmem_ld32(d_acc + lane_base + c0, w);
      tc_wait_ld();
      for (int i = 0; i < 32; ++i) D[tid * kNMax + c0 + i] = __uint_as_float(w[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, 256);
}

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

int main() {
  std::vector<float> A(kRows * kKTot), B(kNMax * kKTot), D(kRows * kNMax);
  for (int r = 0; r < kRows; ++r)
    for (int k = 0; k < kKTot; ++k) A[r * kKTot + k] = static_cast<float>((r * 7 + k * 3) % 11 - 5);
  for (int c = 0; c < kNMax; ++c)
    for (int k = 0; k < kKTot; ++k) B[c * kKTot + k] = static_cast<float>((c * 5 + k * 2) % 13 - 6);
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  const int smem = sizeof(ProbeSmem);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int failures = 0;
  auto run = [&](int mode, int n, const char* name, int k_used) -> int {
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xFF, D.size() * 4);
    probe_kernel<<<1, 128, smem>>>(mode, n, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-44s CUDA error: %s\n", name, cudaGetErrorString(e));
      return -1;
    }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    if (mode == 0) {
      for (int t = 0; t < 128; ++t)
        for (int i = 0; i < 32; ++i) bad += D[t * 32 + i] != static_cast<float>(t * 1000 + i);
    } else {
      for (int r = 0; r < kRows; ++r)
        for (int c = 0; c < n; ++c) {
          float want = 0.0f;
          for (int k = 0; k < k_used; ++k) want += A[r * kKTot + k] * B[c * kKTot + k];
          if (D[r * kNMax + c] != want) {
            if (bad < 4) printf("   (%d,%d): got %g want %g\n", r, c, D[r * kNMax + c], want);
            ++bad;
          }
        }
    }
    printf("%-44s %s (%d mismatches)\n", name, bad ? "FAIL" : "ok", bad);
    return bad;
  };
  int r;
  if ((r = run(0, 32, "TMEM st/ld round trip", 0)) != 0) failures++;
  if (r < 0) return 1;
  if ((r = run(1, 32, "SS  M128 N32  K8", 8)) != 0) failures++;
  if (r < 0) return 1;
  if ((r = run(1, 160, "SS  M128 N160 K8", 8)) != 0) failures++;
  if (r < 0) return 1;
  if ((r = run(2, 32, "TS  M128 N32  K8 (A from TMEM)", 8)) != 0) failures++;
  if (r < 0) return 1;
  if ((r = run(3, 160, "SS  M128 N160 K16 (two K-steps)", 16)) != 0) failures++;
  if (r < 0) return 1;
  if ((r = run(3, 32, "SS  M128 N32  K16 (two K-steps)", 16)) != 0) failures++;
  if (r < 0) return 1;

  // operand conversion: A[0][0] = 1 + 2^-11 + 2^-12 (needs 12 mantissa bits), B[c][0] = 1, everything else 0
  std::vector<float> A2(A.size(), 0.0f), B2(B.size(), 0.0f);
  const float probe = 1.0f + 1.0f / 2048 + 1.0f / 4096;
  A2[0] = probe;
  for (int c = 0; c < kNMax; ++c) B2[c * kKTot] = 1.0f;
  cudaMemcpy(dA, A2.data(), A2.size() * 4, cudaMemcpyHostToDevice);
  B = B2;
  cudaMemset(dD, 0, D.size() * 4);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  probe_kernel<<<1, 128, smem>>>(1, 32, dA, dB, dD);
  if (cudaDeviceSynchronize() == cudaSuccess) {
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    const float got = D[0];
    const char* how = got == 1.0f ? "TRUNCATED to 10 mantissa bits"
                      : got == 1.0f + 1.0f / 1024 ? "ROUNDED to nearest TF32"
                      : got == probe ? "used with ALL 23 mantissa bits" : "something else";
    printf("operand 1 + 2^-11 + 2^-12 times 1.0 = %.10f  -> 32-bit operands are %s\n", got, how);
  }
  printf(failures ? "PROBE FAILED (%d steps)\n" : "PROBE OK\n", failures);
  return failures ? 1 : 0;
}
