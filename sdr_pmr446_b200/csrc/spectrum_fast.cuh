// wf_accumulate_fast_kernel: the asgram / spgram periodogram (/root/reference/src/sdr_pmr446.c:473-477, :910-913; SURVEY.md
// Appendix A.13) for widths W = 8 P with a small smooth P (W = 120: P = 15, the 480-point transform of a 120-column
// terminal), one WARP per transform and nothing but registers between the stages.
//
// The transform is zero-padded: X[k] = sum_{n < W} x[n] w[n] e^{-2 pi i n k / 4W}.  With k = 4 q + r this is, for each of
// the four residues r, a W-point DFT of the pre-twiddled window x_r[n] = x[n] w[n] W_4W^{n r} -- four independent
// 8P-point transforms, one per quarter-warp:
//   lane = (r, l): holds x_r[l + 8 m], m < P (the window value and the pre-twiddle are one complex constant per register)
//   1. P-point DFT over m in registers (compile-time mixed radix 2/3/4/5, twiddles from the kernel parameters),
//   2. twiddle W_8P^{l k2} (lane constants),
//   3. 8-point DFT over the quarter-warp's lanes: transposed through 8 x P float2 of shared memory per quarter-warp, each
//      lane then runs one or two 8-point DFTs in registers,
//   4. |X|^2 accumulated in registers across all the transforms the warp walks; written once per block.
// Round 1's warp kernel kept every stage in shared memory (a mixed-radix Stockham pass per factor): 7.6 ms per 1024-stream
// step at W = 120; the generic kernels remain for every other width.
#pragma once
#include <cuda_runtime.h>

#include "audio_fft.cuh"   // cmul
// included from spectrum.cuh (dft3, dft5, WfParams are defined above the include point)

namespace pmr {

constexpr int WFF_MAXP = 20;

struct WfFastParams {
  WfParams w;
  float2 twp[WFF_MAXP];   // W_P^m, m < P
};

constexpr int WFF_WARPS = 4;   // 128 threads, ~165 registers: three blocks = twelve warps per SM

template <int P>
__global__ void __launch_bounds__(32 * WFF_WARPS, 3) wf_accumulate_fast_kernel(WfFastParams fp) {
  const WfParams& p = fp.w;
  constexpr int W = 8 * P;
  constexpr int NC = (P + 7) / 8;                           // DFT8 columns per lane (k2 = l, l + 8, ...)
  constexpr int PS = P | 1;                                 // odd row stride: the eight lanes of a quarter-warp hit eight banks
  __shared__ float2 xch[WFF_WARPS][4][8 * PS + 1];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int r = lane >> 3, l = lane & 7;
  float2* const xq = xch[wp][r];
  const int s = blockIdx.x / p.parts, part = blockIdx.x % p.parts;
  const float2* res = p.res + (long long)s * p.res_stride;
  // lane constants: window x pre-twiddle W_4W^{n r} for n = l + 8 m, and the inter-stage twiddle W_8P^{l k2}
  float2 cw[P], t2[P];
#pragma unroll
  for (int m = 0; m < P; m++) {
    const int n = l + 8 * m;
    const float2 t = p.twiddle[(n * r) % (4 * W)];          // table of W_4W^k
    const float wv = p.window[n];
    cw[m] = make_float2(wv * t.x, wv * t.y);
    t2[m] = p.twiddle[(4 * l * m) % (4 * W)];               // W_8P^{l m} = W_4W^{4 l m}
  }
  float acc[NC][8];
#pragma unroll
  for (int c = 0; c < NC; c++)
#pragma unroll
    for (int k = 0; k < 8; k++) acc[c][k] = 0.0f;

  const int nw = blockDim.x >> 5;
  const unsigned r0 = (unsigned)p.r0, rmask = (unsigned)p.res_mask;
  auto load = [&](int t, float2* x) {
    const int first = t * p.hop - W;                        // chunk-local index of window sample 0 (may be negative: zeros)
#pragma unroll
    for (int m = 0; m < P; m++) {
      const int li = first + l + 8 * m;
      x[m] = li >= 0 ? res[(r0 + (unsigned)li) & rmask] : make_float2(0.0f, 0.0f);
    }
  };
  int t = 1 + part + p.parts * wp;
  const int tstep = p.parts * nw;
  float2 xn[P];
  if (t <= p.n_transforms) load(t, xn);
  for (; t <= p.n_transforms; t += tstep) {
    float2 x[P], y[P];
#pragma unroll
    for (int m = 0; m < P; m++) x[m] = cmul(xn[m], cw[m]);
    if (t + tstep <= p.n_transforms) load(t + tstep, xn);   // next transform's samples in flight during this one
    RegDft<P, 1>::run(x, 1, y, fp.twp);
#pragma unroll
    for (int k2 = 0; k2 < P; k2++) xq[l * PS + k2] = (l == 0 || k2 == 0) ? y[k2] : cmul(y[k2], t2[k2]);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const int k2 = l + 8 * c;
      if (k2 < P) {
        float2 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = xq[j * PS + k2];
        dft8(v);
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) acc[c][k1] = fmaf(v[k1].x, v[k1].x, fmaf(v[k1].y, v[k1].y, acc[c][k1]));
      }
    }
    __syncwarp();
  }
  // block reduction in a fixed order (bit-reproducible): every warp parks its bins -- bin k = 4 q + r with
  // q = k2 + P k1 -- in the idle exchange buffer, then the block adds the warps in order
  __syncthreads();
  float* red = (float*)&xch[0][0][0];                        // [warps][4 W] floats: half of the exchange buffer
#pragma unroll
  for (int c = 0; c < NC; c++) {
    const int k2 = l + 8 * c;
    if (k2 < P) {
#pragma unroll
      for (int k1 = 0; k1 < 8; k1++) red[wp * (4 * W) + 4 * (k2 + P * k1) + r] = acc[c][k1];
    }
  }
  __syncthreads();
  float* out = p.partial + ((long long)s * p.parts + part) * (4 * W);
  for (int i = threadIdx.x; i < 4 * W; i += blockDim.x) {
    float sum = 0.0f;
    for (int q = 0; q < nw; q++) sum += red[q * (4 * W) + i];
    out[i] = sum;
  }
}

}  // namespace pmr
