// Back-end kernels of the dsd_in chain: discriminator on the 12.5 kHz complex stream, real
// interpolating multi-stage resampler (arbitrary x1.92 then one half-band x2), s16 conversion.
// Replaces freqdem_demodulate_block, msresamp_rrrf_execute and the cast loop of
// /root/reference/src/dsd_in.c:169-175 (rows a13, a14).
// The rates here are tiny (12.5 k -> 48 k samples/s per stream) next to the 1-2.4 Msps front end,
// so these are plain one-thread-per-output kernels over the library's rings.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmr {

// fm[n] = arg(conj(x[n-1]) x[n]) * ref, n in [n0, n1); x[-1] = 0
static __global__ void dsd_freqdem_kernel(const float2* res, long long res_stride, long long res_mask, float* fm, long long fm_stride, long long fm_mask,
                                   long long n0, long long n1, float ref) {
  const long long n = n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n1) return;
  const int s = blockIdx.y;
  const float2* r = res + (long long)s * res_stride;
  const float2 x = r[n & res_mask];
  float2 p = make_float2(0.0f, 0.0f);
  if (n > 0) p = r[(n - 1) & res_mask];
  const float re = __fadd_rn(__fmul_rn(p.x, x.x), __fmul_rn(p.y, x.y));
  const float im = __fsub_rn(__fmul_rn(p.x, x.y), __fmul_rn(p.y, x.x));
  fm[(long long)s * fm_stride + (n & fm_mask)] = atan2f(im, re) * ref;
}

// arbitrary resampler on a real ring (A.5): z[k] = sum_t pfb[idx_k][t] * fm[i_k - t],
// i_k = floor(k*step / 2^24), idx_k = (k*step mod 2^24) >> (24 - bits)
static __global__ void dsd_arb_kernel(const float* fm, long long fm_stride, long long fm_mask, float* z, long long z_stride, long long z_mask,
                               long long k0, long long k1, unsigned step, int bits, const float* pfb) {
  const long long k = k0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= k1) return;
  const int s = blockIdx.y;
  const unsigned long long ph = (unsigned long long)k * step;
  const long long i = (long long)(ph >> 24);
  const unsigned idx = (unsigned)(ph & 0xffffffu) >> (24 - bits);
  const float* row = pfb + ((size_t)idx << 4);
  const float* f = fm + (long long)s * fm_stride;
  float acc = 0.0f;
#pragma unroll
  for (int t = 0; t < 14; t++) {
    const long long n = i - t;
    const float v = n >= 0 ? f[n & fm_mask] : 0.0f;
    acc = fmaf(__ldg(row + t), v, acc);
  }
  z[(long long)s * z_stride + (k & z_mask)] = acc;
}

// half-band interpolator (A.4) + s16: out[2k] = z[k - m], out[2k+1] = sum_j h[j] z[k - j]
struct DsdInterpParams {
  const float* z;
  long long z_stride, z_mask;
  long long k0, k1;
  int m;
  float hb[20];
  float* audio;      // optional [n_streams][out_ld]
  short* pcm;        // optional [n_streams][out_ld]
  long long out_ld;
};
static __global__ void dsd_interp_kernel(DsdInterpParams p) {
  const long long k = p.k0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.k1) return;
  const int s = blockIdx.y;
  const float* z = p.z + (long long)s * p.z_stride;
  const long long d = k - p.m;
  const float y0 = d >= 0 ? z[d & p.z_mask] : 0.0f;
  float y1 = 0.0f;
  for (int j = 0; j < 2 * p.m; j++) {
    const long long n = k - j;
    const float v = n >= 0 ? z[n & p.z_mask] : 0.0f;
    y1 = fmaf(p.hb[j], v, y1);
  }
  const long long o = 2 * (k - p.k0);
  if (p.audio) {
    p.audio[(long long)s * p.out_ld + o] = y0;
    p.audio[(long long)s * p.out_ld + o + 1] = y1;
  }
  if (p.pcm) {
    p.pcm[(long long)s * p.out_ld + o] = (short)__float2int_rz(y0 * 32767.0f);
    p.pcm[(long long)s * p.out_ld + o + 1] = (short)__float2int_rz(y1 * 32767.0f);
  }
}

}  // namespace pmr
