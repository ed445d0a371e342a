// probe: does ptxas accept fma.rn.f32x2 for sm_100a and what SASS does it make
#include <cstdio>
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
  unsigned long long da, aa, bb;
  aa = *reinterpret_cast<unsigned long long*>(&a);
  bb = *reinterpret_cast<unsigned long long*>(&b);
  da = *reinterpret_cast<unsigned long long*>(&d);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(da) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&da);
}
__global__ void k(const float2* x, const float2* y, float2* o, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float2 acc[8];
  for (int j = 0; j < 8; ++j) acc[j] = make_float2(0.f, 0.f);
  for (int t = 0; t < n; ++t) {
    float2 a = x[i + t * 1024];
    for (int j = 0; j < 8; ++j) ffma2(acc[j], a, y[j + t * 8]);
  }
  for (int j = 0; j < 8; ++j) o[i * 8 + j] = acc[j];
}
