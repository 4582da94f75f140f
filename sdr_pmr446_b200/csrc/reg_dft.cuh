// Small DFTs held entirely in registers: 3 and 5 points with constant coefficients, 8 points, and a compile-time
// mixed-radix DFT of any size 2^a 3^b 5^c (twiddles from a table of W_P^m).  Shared by the waterfall kernels
// (spectrum.cuh, spectrum_fast.cuh) and the generic channelizer's FFT.
#pragma once
#include <cuda_runtime.h>

#include "audio_fft.cuh"   // cmul, dft4, dft16

namespace pmr {

// forward DFTs of 3 and 5 points with constant coefficients (in place: v[q] <- sum_r v[r] W_R^(q r))
__device__ __forceinline__ void dft3(float2* v) {
  const float s = 0.86602540378443865f;
  const float2 t1 = make_float2(v[1].x + v[2].x, v[1].y + v[2].y);
  const float2 t2 = make_float2(fmaf(-0.5f, t1.x, v[0].x), fmaf(-0.5f, t1.y, v[0].y));
  const float2 t3 = make_float2(s * (v[1].x - v[2].x), s * (v[1].y - v[2].y));
  v[0] = make_float2(v[0].x + t1.x, v[0].y + t1.y);
  v[1] = make_float2(t2.x + t3.y, t2.y - t3.x);   // t2 - i t3
  v[2] = make_float2(t2.x - t3.y, t2.y + t3.x);   // t2 + i t3
}
__device__ __forceinline__ void dft5(float2* v) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f, s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  const float2 a1 = make_float2(v[1].x + v[4].x, v[1].y + v[4].y), a2 = make_float2(v[2].x + v[3].x, v[2].y + v[3].y);
  const float2 b1 = make_float2(v[1].x - v[4].x, v[1].y - v[4].y), b2 = make_float2(v[2].x - v[3].x, v[2].y - v[3].y);
  const float2 p1 = make_float2(fmaf(c2, a2.x, fmaf(c1, a1.x, v[0].x)), fmaf(c2, a2.y, fmaf(c1, a1.y, v[0].y)));
  const float2 p2 = make_float2(fmaf(c1, a2.x, fmaf(c2, a1.x, v[0].x)), fmaf(c1, a2.y, fmaf(c2, a1.y, v[0].y)));
  const float2 q1 = make_float2(fmaf(s2, b2.x, s1 * b1.x), fmaf(s2, b2.y, s1 * b1.y));
  const float2 q2 = make_float2(fmaf(-s1, b2.x, s2 * b1.x), fmaf(-s1, b2.y, s2 * b1.y));
  v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
  v[1] = make_float2(p1.x + q1.y, p1.y - q1.x);   // p1 - i q1
  v[4] = make_float2(p1.x - q1.y, p1.y + q1.x);   // p1 + i q1
  v[2] = make_float2(p2.x + q2.y, p2.y - q2.x);   // p2 - i q2
  v[3] = make_float2(p2.x - q2.y, p2.y + q2.x);   // p2 + i q2
}

// in-register DFT of compile-time size P = 2^a 3^b 5^c (decimation in time): y[k] = sum_n x[n stride] W_P^{n k}
// TW = stride of this level's twiddles in the W_P0 table (W_P^m = table[m * TW])
template <int P, int TW>
struct RegDft {
  static constexpr int R = (P % 4 == 0) ? 4 : (P % 2 == 0) ? 2 : (P % 3 == 0) ? 3 : (P % 5 == 0) ? 5 : P;
  static constexpr int Q = P / R;
  static_assert(R == 2 || R == 3 || R == 4 || R == 5, "P must factor into 2, 3 and 5");
  __device__ __forceinline__ static void run(const float2* x, int stride, float2* y, const float2* tw, int tws = 1) {
    // n = Q n1 + n2, k = k1 + R k2:  X[k1 + R k2] = sum_n2 W_P^{n2 k1} ( sum_n1 x[Q n1 + n2] W_R^{n1 k1} ) W_Q^{n2 k2}
    float2 a[Q][R];
#pragma unroll
    for (int n2 = 0; n2 < Q; n2++) {
      float2 v[R];
#pragma unroll
      for (int n1 = 0; n1 < R; n1++) v[n1] = x[(Q * n1 + n2) * stride];
      if (R == 2) {
        const float2 t = v[0];
        v[0] = make_float2(t.x + v[1].x, t.y + v[1].y);
        v[1] = make_float2(t.x - v[1].x, t.y - v[1].y);
      } else if (R == 4) {
        dft4<false>(v[0], v[1], v[2], v[3]);
      } else if (R == 3) {
        dft3(v);
      } else {
        dft5(v);
      }
#pragma unroll
      for (int k1 = 0; k1 < R; k1++) a[n2][k1] = (n2 * k1 == 0) ? v[k1] : cmul(v[k1], tw[((n2 * k1) % P) * TW * tws]);
    }
    if (Q == 1) {
#pragma unroll
      for (int k1 = 0; k1 < R; k1++) y[k1] = a[0][k1];
    } else {
#pragma unroll
      for (int k1 = 0; k1 < R; k1++) {
        float2 col[Q], out[Q];
#pragma unroll
        for (int n2 = 0; n2 < Q; n2++) col[n2] = a[n2][k1];
        RegDft<Q, TW * R>::run(col, 1, out, tw, tws);
#pragma unroll
        for (int k2 = 0; k2 < Q; k2++) y[k1 + R * k2] = out[k2];
      }
    }
  }
};
template <int TW>
struct RegDft<1, TW> {
  __device__ __forceinline__ static void run(const float2* x, int stride, float2* y, const float2*, int = 1) { y[0] = x[0]; }
};

// forward 8-point DFT in registers, natural order in and out
__device__ __forceinline__ void dft8(float2* v) {
  const float r = 0.70710678118654752f;
  float2 e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
  dft4<false>(e[0], e[1], e[2], e[3]);
  dft4<false>(o[0], o[1], o[2], o[3]);
  // W8^k o[k]: W8 = (1 - i) / sqrt 2
  const float2 o1 = make_float2(r * (o[1].x + o[1].y), r * (o[1].y - o[1].x));
  const float2 o2 = make_float2(o[2].y, -o[2].x);
  const float2 o3 = make_float2(r * (o[3].y - o[3].x), -r * (o[3].x + o[3].y));
  v[0] = make_float2(e[0].x + o[0].x, e[0].y + o[0].y); v[4] = make_float2(e[0].x - o[0].x, e[0].y - o[0].y);
  v[1] = make_float2(e[1].x + o1.x, e[1].y + o1.y);     v[5] = make_float2(e[1].x - o1.x, e[1].y - o1.y);
  v[2] = make_float2(e[2].x + o2.x, e[2].y + o2.y);     v[6] = make_float2(e[2].x - o2.x, e[2].y - o2.y);
  v[3] = make_float2(e[3].x + o3.x, e[3].y + o3.y);     v[7] = make_float2(e[3].x - o3.x, e[3].y - o3.y);
}

}  // namespace pmr
