// Per-instruction sustained issue rates on sm_100a (8 independent chains per thread, full occupancy).
#include <cstdio>
#include <cuda_runtime.h>

enum { FFMA, FFMA2, FMUL, IMAD_LO, IMAD_HI, IMAD_WIDE, LOP3, IADD3, SHF, PRMT, LDS_RAND, LDS128_BCAST, FMNMX, MIX_FFMA_WIDE, MIX_FFMA2_WIDE, MIX_FFMA_LOP, MIX_FFMA2_LOP, MIX_WIDE_LOP, NMODES };
const char* NAMES[] = {"FFMA", "FFMA2", "FMUL", "IMAD (lo)", "IMAD.HI.U32", "IMAD.WIDE.U32", "LOP3", "IADD3", "SHF", "PRMT", "LDS (random 4B)", "LDS.128 (broadcast)", "FMNMX", "FFMA + IMAD.WIDE 1:1", "FFMA2 + IMAD.WIDE 1:1", "FFMA + LOP3 1:1", "FFMA2 + LOP3 1:1", "IMAD.WIDE + LOP3 1:1"};

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(unsigned* out, int n, unsigned seed) {
  __shared__ float4 sm4[256];
  float* sm = reinterpret_cast<float*>(sm4);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 0.5f;
  __syncthreads();
  unsigned u[8]; float f[8]; unsigned long long w[8]; float2 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { u[i] = threadIdx.x * 7 + i + seed; f[i] = u[i] * 1e-3f; w[i] = u[i]; p[i] = make_float2(f[i], f[i] + 1); }
  const float b = 1.0001f, c = 0.5f; const float2 bb = {b, b}, cc = {c, c};
  const unsigned m = 0xD2511F53u + seed;
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == FFMA) f[i] = fmaf(f[i], b, c);
        if (MODE == FFMA2) p[i] = __ffma2_rn(p[i], bb, cc);
        if (MODE == FMUL) f[i] = f[i] * b;
        if (MODE == IMAD_LO) u[i] = u[i] * m + 12345u;
        if (MODE == IMAD_HI) u[i] = __umulhi(u[i], m) + 1u;  // may fuse into IMAD.HI with addend
        if (MODE == IMAD_WIDE) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((unsigned)(w[i] >> 32) ^ (unsigned)w[i]), "r"(m));
        if (MODE == LOP3) u[i] = (u[i] ^ m) & (u[(i + 1) & 7] | 0x55u);
        if (MODE == IADD3) u[i] = u[i] + u[(i + 1) & 7] + m;
        if (MODE == SHF) u[i] = __funnelshift_l(u[i], u[(i + 1) & 7], 7);
        if (MODE == PRMT) u[i] = __byte_perm(u[i], u[(i + 1) & 7], 0x1234);
        if (MODE == LDS_RAND) u[i] = __float_as_uint(sm[(u[i] >> 3) & 1023]) + i;
        if (MODE == LDS128_BCAST) { float4 v = sm4[(u[i] & 0) + ((it + j + i) & 255)]; f[i] += v.x; u[i] = __float_as_uint(v.y + v.z + v.w) & 0; }
        if (MODE == FMNMX) f[i] = fmaxf(f[i], f[(i + 1) & 7] - 1.0f);
        if (MODE == MIX_FFMA_WIDE) { f[i] = fmaf(f[i], b, c); asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((unsigned)(w[i] >> 32) ^ (unsigned)w[i]), "r"(m)); }
        if (MODE == MIX_FFMA2_WIDE) { p[i] = __ffma2_rn(p[i], bb, cc); asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((unsigned)(w[i] >> 32) ^ (unsigned)w[i]), "r"(m)); }
        if (MODE == MIX_FFMA_LOP) { f[i] = fmaf(f[i], b, c); u[i] = (u[i] ^ m) & (u[(i + 1) & 7] | 0x55u); }
        if (MODE == MIX_FFMA2_LOP) { p[i] = __ffma2_rn(p[i], bb, cc); u[i] = (u[i] ^ m) & (u[(i + 1) & 7] | 0x55u); }
        if (MODE == MIX_WIDE_LOP) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(u[i]), "r"(m)); u[i] = ((unsigned)(w[i] >> 32) ^ m) & ((unsigned)w[i] | 0x55u); }
      }
    }
  }
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= u[i] ^ __float_as_uint(f[i]) ^ (unsigned)w[i] ^ (unsigned)(w[i] >> 32) ^ __float_as_uint(p[i].x + p[i].y);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(double instr_per_slot) {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  unsigned* out; cudaMalloc(&out, sms * 1024 * 4);
  const int n = 400;
  k<MODE><<<sms, 1024>>>(out, 4, 1);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<sms, 1024>>>(out, n, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double slots = (double)n * 16 * 8 * 32 * sms;  // warp-level (mode statements) executed
  double cycles = ms * 1e-3 * khz * 1e3;
  printf("%-26s %8.3f ms  %6.3f statements/cycle/SM  (x%.0f instr => %6.3f warp-instr/cycle/SM)\n", NAMES[MODE], ms, slots / cycles / sms, instr_per_slot,
         slots * instr_per_slot / cycles / sms);
  cudaFree(out);
}

int main() {
  run<FFMA>(1); run<FFMA2>(1); run<FMUL>(1); run<IMAD_LO>(1); run<IMAD_HI>(1); run<IMAD_WIDE>(1); run<LOP3>(1); run<IADD3>(1); run<SHF>(1); run<PRMT>(1);
  run<LDS_RAND>(1); run<LDS128_BCAST>(1); run<FMNMX>(1); run<MIX_FFMA_WIDE>(2); run<MIX_FFMA2_WIDE>(2); run<MIX_FFMA_LOP>(2); run<MIX_FFMA2_LOP>(2); run<MIX_WIDE_LOP>(2);
  return 0;
}
