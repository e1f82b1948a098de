// FMUL2 (packed f32x2 multiply) semantics check on the GPU: elementwise, halves fed by separate loads.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const float2* a, const float2* b, float2* o, float* o2) {
  int i = threadIdx.x;
  o[i] = __fmul2_rn(a[i], b[i]);
  // chain like the MC loop: separate scalar loads into pairs
  const float* af = reinterpret_cast<const float*>(a);
  const float* bf = reinterpret_cast<const float*>(b);
  float2 p = make_float2(af[2 * i], bf[2 * i]);
  p = __fmul2_rn(p, make_float2(af[2 * i + 1], bf[2 * i + 1]));
  o2[2 * i] = p.x; o2[2 * i + 1] = p.y;
}
int main() {
  float2 ha[32], hb[32], ho[32]; float ho2[64];
  for (int i = 0; i < 32; ++i) { ha[i] = make_float2(1.0f + i, 2.0f); hb[i] = make_float2(3.0f, 0.5f + i); }
  float2 *a, *b, *o; float* o2;
  cudaMalloc(&a, sizeof ha); cudaMalloc(&b, sizeof hb); cudaMalloc(&o, sizeof ho); cudaMalloc(&o2, sizeof ho2);
  cudaMemcpy(a, ha, sizeof ha, cudaMemcpyHostToDevice); cudaMemcpy(b, hb, sizeof hb, cudaMemcpyHostToDevice);
  k<<<1, 32>>>(a, b, o, o2);
  cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost); cudaMemcpy(ho2, o2, sizeof ho2, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 3; ++i) printf("i=%d a=(%g,%g) b=(%g,%g) -> fmul2=(%g,%g) expect (%g,%g); chain=(%g,%g) expect (%g,%g)\n", i, ha[i].x, ha[i].y, hb[i].x, hb[i].y, ho[i].x, ho[i].y, ha[i].x*hb[i].x, ha[i].y*hb[i].y, ho2[2*i], ho2[2*i+1], ha[i].x*ha[i].y, hb[i].x*hb[i].y);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
