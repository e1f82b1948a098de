// Rates of conversion instructions on sm_100a (candidate replacements for the quarter-rate IMAD.HI index scaling).
#include <cstdio>
#include <cuda_runtime.h>
enum { F2I, I2F, F2I_FMA, MK_FLOAT_FMA_F2I, LEA_ONLY, NM };
const char* NAMES[] = {"F2I.TRUNC (cvt.rzi.u32.f32)", "I2F (cvt.rn.f32.u32)", "FFMA + F2I", "LOP3 + FFMA + F2I (float index)", "LEA (shl+add)"};
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(unsigned* out, int n, float nf) {
  unsigned u[8]; float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { u[i] = threadIdx.x * 977u + i * 131u; f[i] = threadIdx.x * 0.37f + i; }
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == F2I) { u[i] = __float2uint_rz(f[i]); f[i] = __uint_as_float((u[i] & 0xff) | 0x42000000u); }
        if (MODE == I2F) { f[i] = __uint2float_rn(u[i]); u[i] = __float_as_uint(f[i]) >> 3; }
        if (MODE == F2I_FMA) { u[i] = __float2uint_rz(fmaf(f[i], nf, -nf)); f[i] = __uint_as_float((u[i] & 0x7fffff) | 0x3f800000u); }
        if (MODE == MK_FLOAT_FMA_F2I) { float x = __uint_as_float((u[i] >> 9) | 0x3f800000u); u[i] = u[i] * 2654435761u + __float2uint_rz(fmaf(x, nf, -nf)); }
        if (MODE == LEA_ONLY) { u[i] = (u[i] << 2) + u[(i + 1) & 7]; }
      }
    }
  }
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= u[i] ^ __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  unsigned* out; cudaMalloc(&out, sms * 1024 * 4);
  const int n = 200;
  k<MODE><<<sms, 1024>>>(out, 4, 50.f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<sms, 1024>>>(out, n, 50.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double stm = (double)n * 16 * 8 * 32 * sms, cycles = ms * 1e-3 * khz * 1e3;
  printf("%-36s %.3f ms  %.3f statements/cycle/SM\n", NAMES[MODE], ms, stm / cycles / sms);
  cudaFree(out);
}
int main() { run<F2I>(); run<I2F>(); run<F2I_FMA>(); run<MK_FLOAT_FMA_F2I>(); run<LEA_ONLY>(); return 0; }
