// Can scalar FFMA (fmalite) run concurrently with packed FFMA2 (fmaheavy)?  FMA lanes per cycle per SM for mixes.
#include <cstdio>
#include <cuda_runtime.h>

template <int N2, int N1>   // N2 FFMA2 + N1 FFMA per inner statement group, 8 independent chains each
__global__ void __launch_bounds__(1024, 1) k(float* out, int n) {
  float2 p[8]; float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { p[i] = make_float2(threadIdx.x + i, threadIdx.x - i); f[i] = threadIdx.x * 0.5f + i; }
  const float b = 1.0001f, c = 0.5f; const float2 bb = {b, b}, cc = {c, c};
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int u = 0; u < N2; ++u) p[i] = __ffma2_rn(p[i], bb, cc);
#pragma unroll
        for (int u = 0; u < N1; ++u) f[i] = fmaf(f[i], b, c);
      }
    }
  }
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += p[i].x + p[i].y + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int N2, int N1>
void run() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sms * 1024 * 4);
  const int n = 300;
  k<N2, N1><<<sms, 1024>>>(out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<N2, N1><<<sms, 1024>>>(out, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double groups = (double)n * 8 * 8 * 32 * sms;       // warp-level statement groups
  const double cycles = ms * 1e-3 * khz * 1e3;
  const double fma_lanes = groups * 32.0 * (2.0 * N2 + N1) / cycles / sms;
  printf("FFMA2:FFMA = %d:%d   %.3f ms   %.2f warp-instr/cycle/SM   %.1f FMA lanes/cycle/SM\n", N2, N1, ms,
         groups * (N2 + N1) / cycles / sms, fma_lanes);
  cudaFree(out);
}

int main() {
  run<1, 0>(); run<0, 1>(); run<1, 1>(); run<2, 1>(); run<1, 2>(); run<3, 1>(); run<1, 3>(); run<4, 1>();
  return 0;
}
