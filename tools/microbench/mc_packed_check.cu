// Unit check on the GPU: mc_rounds_xn<20, true, 2> (two chains, packed FMUL2 products) against two single-chain mc_rounds calls
// on the same q table and generators -- must be bit-identical.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I m6anet_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "m6a_mc.cuh"
using namespace m6a;
__global__ void k(int n, int r0, int r1, float* out) {
  __shared__ float q[256];
  const int lane = threadIdx.x;
  for (int i = lane; i < 256; i += 32) q[i] = 0.90f + 0.0003f * static_cast<float>((i * 37) % 251);
  __syncwarp();
  uint32_t qa[2] = {mc_smem_u32(q), mc_smem_u32(q)}, nu[2] = {static_cast<uint32_t>(n), static_cast<uint32_t>(n)};
  int rounds[2] = {r0 - (lane > 20 ? 1 : 0), r1 - (lane > 7 ? 2 : 0)};
  Mwc64x g[2], h[2];
  float v[2] = {0.f, 0.f}, w[2] = {0.f, 0.f};
  for (int i = 0; i < 2; ++i) { g[i].seed(lane, i, 12345, 99); h[i] = g[i]; }
  mc_rounds_xn<20, true, 2>(qa, nu, g, rounds, v);
  for (int i = 0; i < 2; ++i) mc_rounds<20, true>(qa[i], nu[i], h[i], 0, rounds[i], w[i]);
  out[lane * 4 + 0] = v[0]; out[lane * 4 + 1] = v[1]; out[lane * 4 + 2] = w[0]; out[lane * 4 + 3] = w[1];
}
int main() {
  float* d; cudaMalloc(&d, 32 * 4 * sizeof(float));
  float h[128];
  int bad = 0;
  for (int n : {20, 50, 256}) {
    k<<<1, 32>>>(n, 8, 8, d);
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    for (int l = 0; l < 32; ++l) if (h[l*4] != h[l*4+2] || h[l*4+1] != h[l*4+3]) { if (bad < 6) printf("n=%d lane %d: packed (%.9g, %.9g) single (%.9g, %.9g)\n", n, l, h[l*4], h[l*4+1], h[l*4+2], h[l*4+3]); ++bad; }
  }
  printf("%s; mismatches %d\n", cudaGetErrorString(cudaDeviceSynchronize()), bad);
  return bad != 0;
}
