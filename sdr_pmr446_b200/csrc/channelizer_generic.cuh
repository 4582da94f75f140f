// Generic-M analysis channelizer (any channel count <= 4096 -- large prime factors run an O(p) per output butterfly --, any prototype
// semi-length): the wideband path of BASELINE config 4 (1600 channels of 12.5 kHz from a 20 Msps
// capture).  Same arithmetic as channelize16_kernel (replacing
// /root/reference/src/sdr_pmr446.c:804-823 and :881 for every channel) with M as a parameter:
//   nco_mix_kernel         y[j] = x[j] conj(exp(j theta_j)), theta_j = j dtheta mod 2^32   (once per sample)
//   channelize_generic_kernel  one block = one stream x 8 consecutive frames: per frame the M branch
//                          dot products (taps h[i + nM], samples read coalesced from the mixed ring),
//                          an M-point forward FFT in shared memory (Stockham, radices 4/2/3/5/prime,
//                          shared with the waterfall), and the discriminator against the previous
//                          frame; the 8 frames are staged in shared memory so every channel row gets one
//                          32-byte store.
// Throughput here is bounded by L2 traffic for the tap reads (each sample is used by 2m frames); the
// 16-channel PMR kernel keeps its windows in registers instead.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "spectrum.cuh"

namespace pmr {

static __global__ void nco_mix_kernel(const float2* res, float2* mixed, long long stride, long long mask, long long j0, long long count, unsigned dtheta) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const long long j = j0 + k;
  const float2 v = res[blockIdx.y * stride + (j & mask)];
  const unsigned th = (unsigned)j * dtheta;
  float sn, cs;
  sincospif((float)(int)th * (1.0f / 2147483648.0f), &sn, &cs);
  mixed[blockIdx.y * stride + (j & mask)] = make_float2(fmaf(v.x, cs, v.y * sn), fmaf(v.y, cs, -v.x * sn));
}

struct ChanGenParams {
  const float2* mixed;     // mixed resampler output ring [n_streams][stride]
  long long stride, mask;
  long long r1;            // samples available
  int n_streams, M, p;     // channels, taps per branch (2m)
  int tiles;               // tiles (of CG_FT frames) per stream
  long long tile0;
  long long f0, f1;        // owned frames
  float ref;
  const float* taps;       // [p][M]: tap n (newest first) of every branch, contiguous over the branches
  const float2* twiddle;   // [M]
  int n_stages;
  int radix[16];
  float* demod;            // ring [n_streams*M][demod_stride]
  long long demod_stride, demod_mask;
  float2* chan;            // optional [n_streams][M][chan_ld]
  long long chan_ld;
};

constexpr int CG_FT = 8;   // frames per tile (one 32-byte store per channel row)

static __global__ void __launch_bounds__(256) channelize_generic_kernel(ChanGenParams p) {
  extern __shared__ float2 cg_smem[];
  const int M = p.M;
  float2* A = cg_smem;            // FFT ping
  float2* B = A + M;              // FFT pong
  float2* prev = B + M;           // previous frame per channel
  float* outb = (float*)(prev + M);   // [M][CG_FT] discriminator outputs of the tile
  const int s = blockIdx.x / p.tiles;
  const long long tile = p.tile0 + blockIdx.x % p.tiles;
  const long long fa = tile * CG_FT;
  const float2* x = p.mixed + (long long)s * p.stride;
  for (int c = threadIdx.x; c < M; c += blockDim.x) prev[c] = make_float2(0.0f, 0.0f);
  // frame fa - 1 is recomputed to have the discriminator's previous sample
  for (int ff = -1; ff < CG_FT; ff++) {
    const long long f = fa + ff;
    __syncthreads();
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
      // commutator: sample M f + (M - 1 - i) goes to branch i; X[M - 1 - i] = dot(branch i)
      float ar = 0.0f, ai = 0.0f;
      const float* h = p.taps + i;
      for (int n = 0; n < p.p; n++) {
        const long long j = (long long)M * (f - n) + (M - 1 - i);
        if (j >= 0 && j < p.r1) {
          const float2 v = x[j & p.mask];
          const float t = __ldg(h + (size_t)n * M);
          ar = fmaf(t, v.x, ar);
          ai = fmaf(t, v.y, ai);
        }
      }
      A[M - 1 - i] = make_float2(ar, ai);
    }
    __syncthreads();
    float2 *xa = A, *xb = B;
    int Ns = 1;
    for (int st = 0; st < p.n_stages; st++) {
      const int R = p.radix[st];
      if (R == 4) wf_stage<4>(xa, xb, p.twiddle, M, Ns);
      else if (R == 2) wf_stage<2>(xa, xb, p.twiddle, M, Ns);
      else if (R == 3) wf_stage<3>(xa, xb, p.twiddle, M, Ns);
      else if (R == 5) wf_stage<5>(xa, xb, p.twiddle, M, Ns);
      else wf_stage_generic(R, xa, xb, p.twiddle, M, Ns, threadIdx.x, blockDim.x);
      Ns *= R;
      float2* tmp = xa; xa = xb; xb = tmp;
      __syncthreads();
    }
    for (int c = threadIdx.x; c < M; c += blockDim.x) {
      const float2 y = xa[c];
      float2 pv = prev[c];
      if (f == 0) pv = make_float2(0.0f, 0.0f);
      prev[c] = y;
      if (ff >= 0) {
        const float re = __fadd_rn(__fmul_rn(pv.x, y.x), __fmul_rn(pv.y, y.y));
        const float im = __fsub_rn(__fmul_rn(pv.x, y.y), __fmul_rn(pv.y, y.x));
        outb[c * CG_FT + ff] = atan2f(im, re) * p.ref;
        if (p.chan && f >= p.f0 && f < p.f1) p.chan[((long long)s * M + c) * p.chan_ld + (f - p.f0)] = y;
      }
    }
  }
  __syncthreads();
  // one row segment of CG_FT frames per channel; fa is a multiple of 8, so full tiles are one 32-byte store
  for (int c = threadIdx.x; c < M; c += blockDim.x) {
    float* row = p.demod + ((long long)s * M + c) * p.demod_stride;
    if (fa >= p.f0 && fa + CG_FT <= p.f1) {
      float4* d = (float4*)(row + (fa & p.demod_mask));
      const float* o = outb + c * CG_FT;
      d[0] = make_float4(o[0], o[1], o[2], o[3]);
      d[1] = make_float4(o[4], o[5], o[6], o[7]);
    } else {
      for (int ff = 0; ff < CG_FT; ff++)
        if (fa + ff >= p.f0 && fa + ff < p.f1) row[(fa + ff) & p.demod_mask] = outb[c * CG_FT + ff];
    }
  }
}

// ---- faster variant for prototypes with PT taps per branch when 10 frames of M samples fit in shared memory -------------
// Same tile (stream x 8 frames + the frame before), but the branch filters of all nine frames are computed together:
// a thread takes branch i, loads the 8 + PT consecutive frames' samples of that branch once (coalesced over i) and its
// PT taps, and produces nine dot products from registers (PT x 9 FFMA pairs per 8 + 2 PT loads instead of 2 per load).
// The nine M-point FFTs then run in place in shared memory one after the other, and the discriminator reads consecutive
// frames from there.
template <int PT>
static __global__ void __launch_bounds__(512) channelize_generic_tile_kernel(ChanGenParams p) {
  extern __shared__ float2 cg_smem[];
  const int M = p.M;
  constexpr int NF = CG_FT + 1;                 // frames fa - 1 .. fa + CG_FT - 1
  float2* D = cg_smem;                          // [NF][M] branch outputs, then channel outputs (in place)
  float2* P = D + (size_t)NF * M;               // FFT pong
  float* outb = (float*)(P + M);                // [M][CG_FT] discriminator outputs of the tile
  const int s = blockIdx.x / p.tiles;
  const long long tile = p.tile0 + blockIdx.x % p.tiles;
  const long long fa = tile * CG_FT;
  const float2* x = p.mixed + (long long)s * p.stride;
  // ---- branch filters: sample M f + (M - 1 - i) goes to branch i; X[M - 1 - i] = dot(branch i) ----
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    float2 w[NF + PT - 1];                      // w[q] = sample of frame fa - PT + q
#pragma unroll
    for (int q = 0; q < NF + PT - 1; q++) {
      const long long j = (long long)M * (fa - PT + q) + (M - 1 - i);
      w[q] = (j >= 0 && j < p.r1) ? x[j & p.mask] : make_float2(0.0f, 0.0f);
    }
    float t[PT];
#pragma unroll
    for (int n = 0; n < PT; n++) t[n] = __ldg(p.taps + (size_t)n * M + i);
#pragma unroll
    for (int f9 = 0; f9 < NF; f9++) {           // frame fa - 1 + f9 = w index PT - 1 + f9
      float ar = 0.0f, ai = 0.0f;
#pragma unroll
      for (int n = 0; n < PT; n++) {
        ar = fmaf(t[n], w[PT - 1 + f9 - n].x, ar);
        ai = fmaf(t[n], w[PT - 1 + f9 - n].y, ai);
      }
      D[(size_t)f9 * M + (M - 1 - i)] = make_float2(ar, ai);
    }
  }
  __syncthreads();
  // ---- M-point forward FFT of every frame, result back in D ----
  for (int f9 = 0; f9 < NF; f9++) {
    float2 *xa = D + (size_t)f9 * M, *xb = P;
    int Ns = 1;
    for (int st = 0; st < p.n_stages; st++) {
      const int R = p.radix[st];
      if (R == 4) wf_stage<4>(xa, xb, p.twiddle, M, Ns);
      else if (R == 2) wf_stage<2>(xa, xb, p.twiddle, M, Ns);
      else if (R == 3) wf_stage<3>(xa, xb, p.twiddle, M, Ns);
      else if (R == 5) wf_stage<5>(xa, xb, p.twiddle, M, Ns);
      else wf_stage_generic(R, xa, xb, p.twiddle, M, Ns, threadIdx.x, blockDim.x);
      Ns *= R;
      float2* tmp = xa; xa = xb; xb = tmp;
      __syncthreads();
    }
    if (xa == P) {                              // odd number of stages: bring the result home
      for (int c = threadIdx.x; c < M; c += blockDim.x) D[(size_t)f9 * M + c] = P[c];
      __syncthreads();
    }
  }
  // ---- discriminator against the previous frame, channel output ----
  for (int c = threadIdx.x; c < M; c += blockDim.x) {
    float2 pv = D[c];
#pragma unroll
    for (int ff = 0; ff < CG_FT; ff++) {
      const long long f = fa + ff;
      const float2 y = D[(size_t)(ff + 1) * M + c];
      if (f == 0) pv = make_float2(0.0f, 0.0f);
      const float re = __fadd_rn(__fmul_rn(pv.x, y.x), __fmul_rn(pv.y, y.y));
      const float im = __fsub_rn(__fmul_rn(pv.x, y.y), __fmul_rn(pv.y, y.x));
      outb[c * CG_FT + ff] = atan2f(im, re) * p.ref;
      if (p.chan && f >= p.f0 && f < p.f1) p.chan[((long long)s * M + c) * p.chan_ld + (f - p.f0)] = y;
      pv = y;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < M; c += blockDim.x) {
    float* row = p.demod + ((long long)s * M + c) * p.demod_stride;
    if (fa >= p.f0 && fa + CG_FT <= p.f1) {
      float4* d = (float4*)(row + (fa & p.demod_mask));
      const float* o = outb + c * CG_FT;
      d[0] = make_float4(o[0], o[1], o[2], o[3]);
      d[1] = make_float4(o[4], o[5], o[6], o[7]);
    } else {
      for (int ff = 0; ff < CG_FT; ff++)
        if (fa + ff >= p.f0 && fa + ff < p.f1) row[(fa + ff) & p.demod_mask] = outb[c * CG_FT + ff];
    }
  }
}

inline size_t channelize_generic_tile_smem(int M) { return (size_t)M * ((CG_FT + 2) * sizeof(float2) + CG_FT * sizeof(float)); }

}  // namespace pmr
