// Packed FP32 of sm_100 (Blackwell): fma / add / mul .f32x2 on 64-bit register pairs -> SASS FFMA2 / FADD2 / FMUL2.  One
// instruction does the real AND imaginary part of a complex sample; a scalar tap or twiddle component is a
// scalar-broadcast operand (R.F32 / UR.F32), so no duplicated register is needed.  Measured on B200
// (tools/ffma2_probe.cu, tools/ffma2_latency_probe.cu; profiles/ffma2_*.jsonl): same flop rate as FFMA (74 TFLOP/s),
// latency 4 cycles, issue interval 2 cycles -- i.e. half the issue slots of the scalar form for the same work.
#pragma once
#include <cuda_runtime.h>

namespace pmr {

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*(unsigned long long*)&d) : "l"(*(unsigned long long*)&a), "l"(*(unsigned long long*)&b), "l"(*(unsigned long long*)&c));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*(unsigned long long*)&d) : "l"(*(unsigned long long*)&a), "l"(*(unsigned long long*)&b));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*(unsigned long long*)&d) : "l"(*(unsigned long long*)&a), "l"(*(unsigned long long*)&b));
  return d;
}
// scalar tap times a pair: ptxas emits the scalar-broadcast operand form (R.F32 / UR.F32), no duplicate register
__device__ __forceinline__ float2 fma_tap(float h, float2 x, float2 acc) { return ffma2(make_float2(h, h), x, acc); }

__device__ __forceinline__ float2 fsub2(float2 a, float2 b) { return fadd2(a, make_float2(-b.x, -b.y)); }   // the negation is an operand modifier

}  // namespace pmr
