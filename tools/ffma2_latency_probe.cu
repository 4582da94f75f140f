// Micro-benchmark (B200): dependent-issue latency of FFMA2 vs FFMA.  One warp per SM sub-partition (148 blocks x 128 threads),
// NCH independent accumulator chains per thread: cycles per instruction = max(latency / NCH, issue interval).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lat tools/ffma2_latency_probe.cu && /tmp/lat
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*(unsigned long long*)&d) : "l"(*(unsigned long long*)&a), "l"(*(unsigned long long*)&b), "l"(*(unsigned long long*)&c));
  return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
template <int NCH, bool PACKED, bool VIA_C = false>
__global__ void __launch_bounds__(128) lat(float* out, long long* cyc, int iters, float a, float b) {
  float2 x[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) x[i] = make_float2((float)(threadIdx.x + i), (float)i);
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 64 / NCH; u++) {
#pragma unroll
      for (int i = 0; i < NCH; i++) {
        if (PACKED && VIA_C) x[i] = ffma2(aa, bb, x[i]);   // dependency through the addend, as in a FIR accumulation
        else if (PACKED) x[i] = ffma2(x[i], aa, bb);
        else x[i].x = ffma1(x[i].x, a, b);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += x[i].x + x[i].y;
  if (s == 1234.5f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NCH, bool PACKED, bool VIA_C = false>
void run() {
  float* d; long long* c; cudaMalloc(&d, 4); cudaMalloc(&c, 8);
  const int iters = 4096;
  lat<NCH, PACKED, VIA_C><<<148, 128>>>(d, c, iters, 0.999f, 1e-9f);
  lat<NCH, PACKED, VIA_C><<<148, 128>>>(d, c, iters, 0.999f, 1e-9f);
  long long h = 0; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("{\"op\": \"%s%s\", \"chains\": %d, \"cycles_per_instr\": %.2f}\n", PACKED ? "FFMA2" : "FFMA", VIA_C ? " (chain through addend)" : "", NCH, (double)h / ((double)iters * 64));
  cudaFree(d); cudaFree(c);
}
int main() {
  run<1, false>(); run<2, false>(); run<4, false>(); run<8, false>();
  run<1, true>(); run<2, true>(); run<4, true>(); run<8, true>(); run<16, true>();
  run<1, true, true>(); run<2, true, true>(); run<3, true, true>(); run<4, true, true>(); run<8, true, true>();
  return cudaDeviceSynchronize() != cudaSuccess;
}
