// tcgen05.mma kind::tf32 issue / execution rate for the shapes of the read encoder (one CTA, one issuing thread).
// ELECT=1 (argv[1]): the whole warp runs the loop and one elect.sync lane issues (back-to-back UTCHMMA); ELECT=0: the round-2
// first measurement, `if (tid == 0)`, where ptxas wraps every MMA in a lane loop -- its ">= 46 cycles per MMA" was that loop.
// For every pattern: R back-to-back MMAs, one commit, wait; cycles per MMA = (t_done - t_start) / R, and the cycles the
// issuing thread spent before the commit (issue cost / queue back-pressure).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I m6anet_b200/csrc -o tools/microbench/tcgen05_rate \
//        tools/microbench/tcgen05_rate.cu && tools/microbench/tcgen05_rate
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "m6a_tc.cuh"
using namespace m6a::tc;

struct Smem {
  alignas(128) float a[4][128][4];       // A operand, K = 16
  alignas(128) float b[4][256][4];       // B operand up to N = 256, K = 16
  uint32_t tmem_base;
  alignas(8) unsigned long long bar;
};

// mode: 0 SS, 1 TS.  n: MMA N.  pattern 2 = the Linear-2 chunk pattern (TS N64 + TS N32 alternating), n ignored
template <bool kElect>
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int n, int reps, long long* out) {
  extern __shared__ __align__(128) unsigned char raw[];
  Smem& sm = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4 * 128 * 4; i += 128) (&sm.a[0][0][0])[i] = 1.0f;
  for (int i = tid; i < 4 * 256 * 4; i += 128) (&sm.b[0][0][0])[i] = 0.5f;
  if (tid == 0) { mbar_init(&sm.bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
  fence_proxy_async(); fence_before(); __syncthreads(); fence_after();
  const uint32_t tmem = sm.tmem_base;
  if (kElect ? warp == 0 : tid == 0) {
    const uint32_t sa = smem_u32(sm.a), sb = smem_u32(sm.b);
    const uint64_t da = make_desc(sa, 128 * 16, 128);
    const uint32_t d = tmem, a_tm = tmem + 384;
    long long t0 = clock64();
    const bool issuer = kElect ? elect_one() : true;
    if (!issuer) {
    } else if (mode == 2) {
      const uint64_t db = make_desc(sb, 64 * 16, 128);
      const uint32_t i64 = make_idesc(128, 64), i32 = make_idesc(128, 32);
      for (int r = 0; r < reps; r += 2) { mma_ts(d, a_tm, db, i64, 1u); mma_ts(d + 32, a_tm + 8, db, i32, 1u); }
    } else {
      const uint64_t db = make_desc(sb, n * 16, 128);
      const uint32_t id = make_idesc(128, n);
      if (mode == 0) for (int r = 0; r < reps; ++r) mma_ss(d, da, db, id, 1u);
      else for (int r = 0; r < reps; ++r) mma_ts(d, a_tm, db, id, 1u);
    }
    long long t1 = clock64();
    if (issuer) mma_commit(&sm.bar);
    __syncwarp();
    mbar_wait(&sm.bar, 0, 99);
    long long t2 = clock64();
    if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  fence_before(); __syncthreads();
  if (warp == 0) tmem_free(tmem, 512);
}

int main(int argc, char** argv) {
  const bool elect = argc > 1 && argv[1][0] == '1';
  printf("issue mode: %s\n", elect ? "elect.sync lane of a converged warp" : "if (tid == 0)");
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  auto run = [&](const char* name, int mode, int n, int reps) {
    long long h[2] = {0, 0};
    for (int it = 0; it < 2; ++it) {
      if (elect) rate_kernel<true><<<1, 128, sizeof(Smem)>>>(mode, n, reps, d);
      else rate_kernel<false><<<1, 128, sizeof(Smem)>>>(mode, n, reps, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    }
    printf("%-34s reps %4d: issue %7.1f cyc/MMA   issue+complete %7.1f cyc/MMA (total %lld)\n", name, reps, (double)h[0] / reps,
           (double)h[1] / reps, h[1]);
  };
  for (int reps : {8, 64, 512}) {
    run("SS M128 N32  K8", 0, 32, reps);
    run("SS M128 N64  K8", 0, 64, reps);
    run("SS M128 N160 K8", 0, 160, reps);
    run("SS M128 N256 K8", 0, 256, reps);
    run("TS M128 N32  K8", 1, 32, reps);
    run("TS M128 N64  K8", 1, 64, reps);
    run("TS M128 N128 K8", 1, 128, reps);
    run("TS M128 N256 K8", 1, 256, reps);
    run("TS chunk pattern N64+N32", 2, 0, reps);
  }
  return 0;
}
