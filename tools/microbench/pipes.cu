// Microbenchmarks of sm_100a issue/pipe rates that drive the kernel design (run under gpurun).
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#define ITERS 4096
template <int MODE>
__global__ void k(float* out, unsigned* outi, int n) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  float b = 1.0001f, c = 0.5f;
  float2 p0 = {a0, a1}, p1 = {a2, a3}, p2 = {a4, a5}, p3 = {a6, a7}, p4 = {a1, a2}, p5 = {a3, a4}, p6 = {a5, a6}, p7 = {a7, a0};
  float2 bb = {b, b}, cc = {c, c};
  unsigned u0 = threadIdx.x, u1 = u0 * 3, u2 = u0 * 5, u3 = u0 * 7;
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < ITERS / 64; ++j) {
      if (MODE == 0) {  // 8 independent FFMA chains
        a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
      } else if (MODE == 1) {  // 8 independent FFMA2 chains (16 FMAs)
        p0 = __ffma2_rn(p0, bb, cc); p1 = __ffma2_rn(p1, bb, cc); p2 = __ffma2_rn(p2, bb, cc); p3 = __ffma2_rn(p3, bb, cc);
        p4 = __ffma2_rn(p4, bb, cc); p5 = __ffma2_rn(p5, bb, cc); p6 = __ffma2_rn(p6, bb, cc); p7 = __ffma2_rn(p7, bb, cc);
      } else if (MODE == 2) {  // IMAD.WIDE-like: mulhi + mullo pairs (philox round core)
        unsigned h0 = __umulhi(0xD2511F53u, u0), l0 = 0xD2511F53u * u0;
        unsigned h1 = __umulhi(0xCD9E8D57u, u2), l1 = 0xCD9E8D57u * u2;
        u0 = h1 ^ u1 ^ 0x9E3779B9u; u1 = l1; u2 = h0 ^ u3 ^ 0xBB67AE85u; u3 = l0;
      } else if (MODE == 3) {  // 4 FFMA + 4 LOP3-ish interleaved (dual pipe)
        a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        u0 = (u0 ^ u1) & u2; u1 = (u1 ^ u2) | u3; u2 = (u2 & u3) ^ u0; u3 = (u3 | u0) ^ u1;
      } else if (MODE == 4) {  // 4 FFMA2 + 4 LOP3
        p0 = __ffma2_rn(p0, bb, cc); p1 = __ffma2_rn(p1, bb, cc); p2 = __ffma2_rn(p2, bb, cc); p3 = __ffma2_rn(p3, bb, cc);
        u0 = (u0 ^ u1) & u2; u1 = (u1 ^ u2) | u3; u2 = (u2 & u3) ^ u0; u3 = (u3 | u0) ^ u1;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y +
                                               p4.x + p4.y + p5.x + p5.y + p6.x + p6.y + p7.x + p7.y;
  outi[blockIdx.x * blockDim.x + threadIdx.x] = u0 ^ u1 ^ u2 ^ u3;
}

template <int MODE>
void run(const char* name, double ops_per_inner, int threads, int blocks_per_sm) {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  float* out; unsigned* outi;
  int blocks = sms * blocks_per_sm;
  cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&outi, blocks * threads * 4);
  int n = 200;
  k<MODE><<<blocks, threads>>>(out, outi, 2);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, outi, n);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inner = (double)n * (ITERS / 64);
  double warp_instr = inner * ops_per_inner * (threads / 32) * blocks;
  double cycles = ms * 1e-3 * khz * 1e3;
  printf("%-28s thr=%4d bps=%d  %.3f ms  %.3f warp-instr/cycle/SM (at max clock %d MHz)\n", name, threads, blocks_per_sm, ms,
         warp_instr / cycles / sms, khz / 1000);
  cudaFree(out); cudaFree(outi);
}

int main() {
  for (int thr : {128, 256, 512, 1024}) {
    run<0>("FFMA x8", 8, thr, 1);
    run<1>("FFMA2 x8 (16 fma)", 8, thr, 1);
    run<2>("philox round (2 IMAD.WIDE+2 LOP3)", 4, thr, 1);
    run<3>("4 FFMA + 4 LOP3", 8, thr, 1);
    run<4>("4 FFMA2 + 4 LOP3", 8, thr, 1);
  }
  // pinned H2D / D2H bandwidth
  size_t bytes = 1ull << 30;
  void *h, *d; cudaMallocHost(&h, bytes); cudaMalloc(&d, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("H2D pinned 1 GiB: %.2f ms  %.1f GB/s\n", ms, bytes / ms / 1e6);
    cudaEventRecord(e0); cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("D2H pinned 1 GiB: %.2f ms  %.1f GB/s\n", ms, bytes / ms / 1e6);
  }
  return 0;
}
