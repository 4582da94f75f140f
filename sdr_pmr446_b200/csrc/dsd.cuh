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
// i_k = floor(k*step / 2^24), idx_k = (k*step mod 2^24) >> (24 - bits).
// One block = 256 consecutive outputs of one stream: their inputs are staged in shared memory, the bank rows (16 floats,
// 64-byte aligned) come from global memory with four 128-bit loads.
constexpr int DSD_XS = 1216;  // inputs staged per block: enough for 1024 outputs at any step >= 2^23 (rate <= 2), 4.9 KB
constexpr int DSD_RB = 4;     // outputs per thread: a block of 256 threads covers 1024 consecutive outputs
static __global__ void __launch_bounds__(256) dsd_arb_kernel(const float* fm, long long fm_stride, long long fm_mask, float* z, long long z_stride,
                                                             long long z_mask, long long k0, long long k1, unsigned step, int bits, const float* pfb) {
  __shared__ float xs[DSD_XS];
  const int s = blockIdx.y;
  const long long kb = k0 + (long long)blockIdx.x * (256 * DSD_RB);
  const long long kend = kb + 256 * DSD_RB < k1 ? kb + 256 * DSD_RB : k1;
  if (kb >= k1) return;
  const float* f = fm + (long long)s * fm_stride;
  const long long i_first = (long long)(((unsigned long long)kb * step) >> 24) - 13;
  const long long i_last = (long long)(((unsigned long long)(kend - 1) * step) >> 24);
  const int count = (int)(i_last - i_first + 1);
  const bool staged = count <= DSD_XS;   // block-uniform
  if (staged) {
    for (int c = threadIdx.x; c < count; c += 256) {
      const long long n = i_first + c;
      xs[c] = n >= 0 ? f[n & fm_mask] : 0.0f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < DSD_RB; r++) {
    const long long k = kb + threadIdx.x + 256 * r;
    if (k >= kend) break;
    const unsigned long long ph = (unsigned long long)k * step;
    const long long i = (long long)(ph >> 24);
    const unsigned idx = (unsigned)(ph & 0xffffffu) >> (24 - bits);
    // the row straight from the (L1 / L2 resident, 16 KB) bank: a x1.92 resampler cycles through 25 of its 256 rows
    const float4* row = (const float4*)(pfb + ((size_t)idx << 4));
    const float4 h0 = __ldg(row), h1 = __ldg(row + 1), h2 = __ldg(row + 2), h3 = __ldg(row + 3);
    const float h[14] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, h2.x, h2.y, h2.z, h2.w, h3.x, h3.y};
    float acc = 0.0f;
    if (staged) {
      const float* x = xs + (int)(i - i_first);
#pragma unroll
      for (int t = 0; t < 14; t++) acc = fmaf(h[t], x[-t], acc);
    } else {
#pragma unroll
      for (int t = 0; t < 14; t++) {
        const long long n = i - t;
        const float v = n >= 0 ? f[n & fm_mask] : 0.0f;
        acc = fmaf(h[t], v, acc);
      }
    }
    z[(long long)s * z_stride + (k & z_mask)] = acc;
  }
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
// one block = 256 consecutive input positions k of one stream, staged with their 2m - 1 predecessors
static __global__ void __launch_bounds__(256) dsd_interp_kernel(DsdInterpParams p) {
  __shared__ float zs[256 + 20];
  const int s = blockIdx.y;
  const long long kb = p.k0 + (long long)blockIdx.x * 256;
  if (kb >= p.k1) return;
  const float* z = p.z + (long long)s * p.z_stride;
  const int hist = 2 * p.m - 1;
  for (int c = threadIdx.x; c < 256 + hist; c += 256) {
    const long long n = kb - hist + c;
    zs[c] = (n >= 0 && n < p.k1) ? z[n & p.z_mask] : 0.0f;
  }
  __syncthreads();
  const long long k = kb + threadIdx.x;
  if (k >= p.k1) return;
  const float* q = zs + threadIdx.x + hist;   // q[-j] = z[k - j]
  const float y0 = q[-p.m];
  float y1 = 0.0f;
  for (int j = 0; j < 2 * p.m; j++) y1 = fmaf(p.hb[j], q[-j], y1);
  const long long o = (long long)s * p.out_ld + 2 * (k - p.k0);
  if (p.audio) {
    p.audio[o] = y0;
    p.audio[o + 1] = y1;
  }
  if (p.pcm) {
    const short a = (short)__float2int_rz(y0 * 32767.0f), b = (short)__float2int_rz(y1 * 32767.0f);
    if (((o | (long long)(uintptr_t)p.pcm >> 1) & 1) == 0) {
      *(short2*)(p.pcm + o) = make_short2(a, b);
    } else {
      p.pcm[o] = a;
      p.pcm[o + 1] = b;
    }
  }
}

}  // namespace pmr
