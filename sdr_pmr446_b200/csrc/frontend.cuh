// Front-end kernels: cu8/cf32 load -> DC blocker -> half-band decimator cascade -> arbitrary
// resampler.  Replaces iirfilt_crcf_execute_block + msresamp_crcf_execute of the reference
// (/root/reference/src/sdr_pmr446.c:795-796, /root/reference/src/dsd_in.c:167-168; algorithm:
// SURVEY.md Appendix A.1-A.5).
//
// B200 design ("segment-sequential"): every thread owns one time segment of one stream and
// runs the cascade over it sample-group by sample-group with all filter windows in registers,
// exactly like the CPU does for a whole stream -- but ~10^5-10^6 segments run concurrently.
// A segment starts H samples early (zero state) so that its owned outputs only depend on real
// samples (the cascade is FIR); the one IIR element, the DC blocker, gets its exact state at
// the segment start from a two-level linear-recurrence scan (dc_local_kernel + dc_scan_kernel).
// Taps are compile-time-indexed constant-bank operands of the FFMAs; no shared memory is used.
// Lanes of a warp read 32-byte (cu8) / 128-byte (cf32) pieces of different segments, i.e. whole
// DRAM sectors / lines, which L1/L2 turn into fully used transactions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmr {

enum { SRC_CU8 = 0, SRC_CF32 = 1, SRC_RING = 2 };

// Per-launch view of the input signal of all streams.
//  two-source (SRC_CU8 / SRC_CF32): samples [hist_base, n0) live in `hist`, [n0, n1) in `cur`.
//  ring (SRC_RING): sample n lives at cur[n & ring_mask]; samples with n < 0 read as zero.
struct SrcView {
  const void* hist;
  const void* cur;
  long long hist_stride;  // bytes between streams
  long long cur_stride;   // bytes between streams
  long long hist_base;    // absolute index of hist[0]; multiple of 16
  long long n0, n1;       // new samples are [n0, n1)
  long long ring_mask;
  int cur_aligned;        // n0 % 16 == 0 and cur/cur_stride 16-byte aligned (two-source only)
};

struct CascadeParams {
  SrcView src;
  int n_streams;
  int nseg;               // segments per stream
  long long seg0;         // absolute start of segment 0 (multiple of the group size)
  int seg_len;            // input samples per segment (multiple of group size and of D)
  int halo;               // warm-up samples before a segment (multiple of the group size)
  long long out0, out1;   // owned half-band outputs [out0, out1)
  float scale;            // 2^-NST, folded into the last stage's outputs
  // DC blocker
  float alpha;
  const float2* v_seg;    // [n_streams][nseg] V at (segment start - halo)
  // arbitrary resampler
  unsigned step;
  int bits;
  const float* pfb;       // [npfb][16] rows padded to 16 floats
  // destination ring (cf32)
  float2* dst;
  long long dst_stride;   // float2 elements between streams
  long long dst_mask;
  // half-band taps of the (up to 4) stages of this launch: [stage][j], j < 2m, newest first.
  // Kernel parameters live in constant bank 0, so these become c[0x0][imm] FFMA operands.
  float hb[4][20];
};

template <int SRC>
struct Loader;

// ---- cu8 two-source --------------------------------------------------------------------
__device__ __forceinline__ float u8f(unsigned w, int k) { return (float)((w >> (8 * k)) & 0xffu); }

template <>
struct Loader<SRC_CU8> {
  // loads G samples starting at absolute q (multiple of G) of stream s, converts with the
  // SoapyRTLSDR rule (u8 - 127.4)/128
  template <int G>
  static __device__ __forceinline__ void load(const SrcView& v, int s, long long q, float* xr, float* xi) {
    const float k = 1.0f / 128.0f, c0 = -127.4f / 128.0f;
    const uint8_t* hist = (const uint8_t*)v.hist + (long long)s * v.hist_stride;
    const uint8_t* cur = (const uint8_t*)v.cur + (long long)s * v.cur_stride;
    if (q >= v.n0 && q + G <= v.n1 && v.cur_aligned) {
      const uint4* p = (const uint4*)(cur + 2 * (q - v.n0));
#pragma unroll
      for (int i = 0; i < G / 8; i++) {
        uint4 w = __ldg(p + i);
        unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          xr[i * 8 + 2 * j] = fmaf(u8f(ww[j], 0), k, c0);
          xi[i * 8 + 2 * j] = fmaf(u8f(ww[j], 1), k, c0);
          xr[i * 8 + 2 * j + 1] = fmaf(u8f(ww[j], 2), k, c0);
          xi[i * 8 + 2 * j + 1] = fmaf(u8f(ww[j], 3), k, c0);
        }
      }
    } else if (q + G <= v.n0 && q >= v.hist_base) {
      const uint4* p = (const uint4*)(hist + 2 * (q - v.hist_base));
#pragma unroll
      for (int i = 0; i < G / 8; i++) {
        uint4 w = p[i];
        unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          xr[i * 8 + 2 * j] = fmaf(u8f(ww[j], 0), k, c0);
          xi[i * 8 + 2 * j] = fmaf(u8f(ww[j], 1), k, c0);
          xr[i * 8 + 2 * j + 1] = fmaf(u8f(ww[j], 2), k, c0);
          xi[i * 8 + 2 * j + 1] = fmaf(u8f(ww[j], 3), k, c0);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < G; i++) {  // static indices keep xr/xi in registers
        long long n = q + i;
        float r = 0.0f, im = 0.0f;
        if (n >= 0 && n < v.n1 && (n >= v.n0 || n >= v.hist_base)) {
          const uint8_t* b = (n >= v.n0) ? cur + 2 * (n - v.n0) : hist + 2 * (n - v.hist_base);
          r = fmaf((float)b[0], k, c0);
          im = fmaf((float)b[1], k, c0);
        }
        xr[i] = r;
        xi[i] = im;
      }
    }
  }
};

// ---- cf32 two-source -------------------------------------------------------------------
template <>
struct Loader<SRC_CF32> {
  template <int G>
  static __device__ __forceinline__ void load(const SrcView& v, int s, long long q, float* xr, float* xi) {
    const float2* hist = (const float2*)((const char*)v.hist + (long long)s * v.hist_stride);
    const float2* cur = (const float2*)((const char*)v.cur + (long long)s * v.cur_stride);
    if (q >= v.n0 && q + G <= v.n1 && v.cur_aligned) {
      const float4* p = (const float4*)(cur + (q - v.n0));
#pragma unroll
      for (int i = 0; i < G / 2; i++) {
        float4 w = __ldg(p + i);
        xr[2 * i] = w.x; xi[2 * i] = w.y; xr[2 * i + 1] = w.z; xi[2 * i + 1] = w.w;
      }
    } else if (q + G <= v.n0 && q >= v.hist_base) {
      const float4* p = (const float4*)(hist + (q - v.hist_base));
#pragma unroll
      for (int i = 0; i < G / 2; i++) {
        float4 w = p[i];
        xr[2 * i] = w.x; xi[2 * i] = w.y; xr[2 * i + 1] = w.z; xi[2 * i + 1] = w.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < G; i++) {  // static indices keep xr/xi in registers
        long long n = q + i;
        float2 w = make_float2(0.0f, 0.0f);
        if (n >= 0 && n < v.n1) {
          if (n >= v.n0) w = cur[n - v.n0];
          else if (n >= v.hist_base) w = hist[n - v.hist_base];
        }
        xr[i] = w.x;
        xi[i] = w.y;
      }
    }
  }
};

// ---- cf32 ring (library-owned intermediate) -----------------------------------------------
template <>
struct Loader<SRC_RING> {
  template <int G>
  static __device__ __forceinline__ void load(const SrcView& v, int s, long long q, float* xr, float* xi) {
    const float2* ring = (const float2*)((const char*)v.cur + (long long)s * v.cur_stride);
    if (q >= 0 && q + G <= v.n1) {
      const float4* p = (const float4*)(ring + (q & v.ring_mask));
#pragma unroll
      for (int i = 0; i < G / 2; i++) {
        float4 w = p[i];
        xr[2 * i] = w.x; xi[2 * i] = w.y; xr[2 * i + 1] = w.z; xi[2 * i + 1] = w.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < G; i++) {  // static indices keep xr/xi in registers
        long long n = q + i;
        float2 w = make_float2(0.0f, 0.0f);
        if (n >= 0 && n < v.n1) w = ring[n & v.ring_mask];
        xr[i] = w.x;
        xi[i] = w.y;
      }
    }
  }
};

// ---- one half-band decimator stage held in registers (A.4) ------------------------------
//   out[o] = x_odd[o - M] + sum_{j<2M} h[j] * x_even[o - j]
// processes B input pairs per call; STAGE selects the row of CascadeParams::hb.
template <int M, int B, int STAGE>
struct HbStage {
  float her[2 * M - 1], hei[2 * M - 1];  // previous even samples, oldest first
  float hor_[M], hoi[M];                 // previous odd samples, oldest first
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < 2 * M - 1; i++) her[i] = hei[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < M; i++) hor_[i] = hoi[i] = 0.0f;
  }
  __device__ __forceinline__ void run(const CascadeParams& p, const float* xr, const float* xi, float* yr, float* yi, float scale) {
    float er[2 * M - 1 + B], ei[2 * M - 1 + B], or_[M + B], oi[M + B];
#pragma unroll
    for (int i = 0; i < 2 * M - 1; i++) { er[i] = her[i]; ei[i] = hei[i]; }
#pragma unroll
    for (int i = 0; i < M; i++) { or_[i] = hor_[i]; oi[i] = hoi[i]; }
#pragma unroll
    for (int b = 0; b < B; b++) {
      er[2 * M - 1 + b] = xr[2 * b]; ei[2 * M - 1 + b] = xi[2 * b];
      or_[M + b] = xr[2 * b + 1];    oi[M + b] = xi[2 * b + 1];
    }
#pragma unroll
    for (int b = 0; b < B; b++) {
      float ar = or_[b], ai = oi[b];
#pragma unroll
      for (int j = 0; j < 2 * M; j++) {
        ar = fmaf(p.hb[STAGE][j], er[2 * M - 1 + b - j], ar);
        ai = fmaf(p.hb[STAGE][j], ei[2 * M - 1 + b - j], ai);
      }
      yr[b] = ar * scale;
      yi[b] = ai * scale;
    }
#pragma unroll
    for (int i = 0; i < 2 * M - 1; i++) { her[i] = er[B + i]; hei[i] = ei[B + i]; }
#pragma unroll
    for (int i = 0; i < M; i++) { hor_[i] = or_[B + i]; hoi[i] = oi[B + i]; }
  }
};
template <int B, int STAGE>
struct HbStage<0, B, STAGE> {  // absent stage
  __device__ __forceinline__ void reset() {}
};

// ---- DC blocker local sums: S_t = sum_j c^(len-1-j) x[P_t + j] over segment t (A.1) -------
struct DcLocalParams {
  SrcView src;
  int n_streams, nseg;
  long long p0;       // P_0 = seg0 - halo
  int seg_len;
  long long end;      // sums stop at `end` (next chunk's P_0)
  float c;            // 1 - alpha
  float2* sums;       // [n_streams][nseg]
};

template <int SRC>
__global__ void __launch_bounds__(128) dc_local_kernel(DcLocalParams p) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)p.n_streams * p.nseg) return;
  int s = (int)(gid / p.nseg), t = (int)(gid % p.nseg);
  long long a = p.p0 + (long long)t * p.seg_len;
  long long b = a + p.seg_len;
  if (b > p.end) b = p.end;
  float sr = 0.0f, si = 0.0f;
  if (a < 0) a = 0;  // P_0 and seg_len are multiples of 16, so a stays group-aligned
  for (long long q = a; q < b; q += 16) {
    float xr[16], xi[16];
    Loader<SRC>::template load<16>(p.src, s, q, xr, xi);
    if (q + 16 <= b) {
#pragma unroll
      for (int i = 0; i < 16; i++) { sr = fmaf(p.c, sr, xr[i]); si = fmaf(p.c, si, xi[i]); }
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++)
        if (q + i < b) { sr = fmaf(p.c, sr, xr[i]); si = fmaf(p.c, si, xi[i]); }
    }
  }
  p.sums[gid] = make_float2(sr, si);
}

struct DcScanParams {
  int n_streams, nseg;
  const float2* sums;
  float2* v_seg;      // out: V at each segment's warm-up start
  float2* v_lag;      // in/out: V at P_0 of this chunk -> V at `end`
  long long p0, end;
  int seg_len;
  float c, decay_full;  // 1 - alpha and c^seg_len
};
static __global__ void dc_scan_kernel(DcScanParams p) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.n_streams) return;
  float2 v = p.v_lag[s];
  for (int t = 0; t < p.nseg; t++) {
    p.v_seg[(long long)s * p.nseg + t] = v;
    long long a = p.p0 + (long long)t * p.seg_len, b = a + p.seg_len;
    if (b > p.end) b = p.end;
    if (a < 0) a = 0;
    const long long len = b - a;
    if (len <= 0) continue;
    const float d = (len == p.seg_len) ? p.decay_full : powf(p.c, (float)len);
    float2 sm = p.sums[(long long)s * p.nseg + t];
    v.x = fmaf(d, v.x, sm.x);
    v.y = fmaf(d, v.y, sm.y);
  }
  p.v_lag[s] = v;
}

// ---- the cascade kernel ----------------------------------------------------------------------
// Stage list MA,MB,MC,MD = semi-lengths in execution order (highest rate first), 0 = absent.
// G = input samples per loop iteration (multiple of 2^NST).
template <int SRC, bool DC, int G, int MA, int MB, int MC, int MD, bool ARB>
__global__ void __launch_bounds__(128) cascade_kernel(CascadeParams p) {
  constexpr int NST = (MA > 0) + (MB > 0) + (MC > 0) + (MD > 0);
  constexpr int D = 1 << NST;
  constexpr int NO = G / D;  // outputs per iteration
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)p.n_streams * p.nseg) return;
  const int s = (int)(gid / p.nseg), t = (int)(gid % p.nseg);
  const long long T0 = p.seg0 + (long long)t * p.seg_len;
  long long i_lo = T0 / D, i_hi = (T0 + p.seg_len) / D;
  if (i_lo < p.out0) i_lo = p.out0;
  if (i_hi > p.out1) i_hi = p.out1;
  if (i_hi <= i_lo) return;

  HbStage<MA, G / 2, 0> sa;
  HbStage<MB, G / 4, 1> sb;
  HbStage<MC, G / 8, 2> sc;
  HbStage<MD, G / 16, 3> sd;
  sa.reset(); sb.reset(); sc.reset(); sd.reset();

  float vr = 0.0f, vi = 0.0f;
  long long q = T0 - p.halo;
  if (DC) {
    if (q > 0) {
      float2 v0 = p.v_seg[gid];
      vr = v0.x; vi = v0.y;
    }
  }
  if (q < 0) q = 0;

  // arbitrary resampler state
  float wr[13 + NO], wi[13 + NO];
  unsigned phase = 0;
  long long j = 0;
  if (ARB) {
#pragma unroll
    for (int i = 0; i < 13 + NO; i++) wr[i] = wi[i] = 0.0f;
    if (i_lo > 0) {
      unsigned long long num = (unsigned long long)i_lo << 24;
      j = (long long)((num + p.step - 1) / p.step);
      phase = (unsigned)((unsigned long long)j * p.step - num);
    }
  }
  float2* dst = p.dst + (long long)s * p.dst_stride;
  const long long q_end = i_hi * D;
  const float scale = p.scale;

  for (; q < q_end; q += G) {
    float xr[G], xi[G];
    Loader<SRC>::template load<G>(p.src, s, q, xr, xi);
    if (DC) {
#pragma unroll
      for (int i = 0; i < G; i++) {
        float yr = fmaf(-p.alpha, vr, xr[i]), yi = fmaf(-p.alpha, vi, xi[i]);
        vr += yr; vi += yi;
        xr[i] = yr; xi[i] = yi;
      }
    }
    // half-band stages; the launch's scale rides on the last present stage
    float ar[G / 2 > 0 ? G / 2 : 1], ai[G / 2 > 0 ? G / 2 : 1];
    float br[G / 4 > 0 ? G / 4 : 1], bi[G / 4 > 0 ? G / 4 : 1];
    float cr[G / 8 > 0 ? G / 8 : 1], ci[G / 8 > 0 ? G / 8 : 1];
    float dr[G / 16 > 0 ? G / 16 : 1], di[G / 16 > 0 ? G / 16 : 1];
    float* outr = xr; float* outi = xi;
    if constexpr (MA > 0) { sa.run(p, xr, xi, ar, ai, NST == 1 ? scale : 1.0f); outr = ar; outi = ai; }
    if constexpr (MB > 0) { sb.run(p, ar, ai, br, bi, NST == 2 ? scale : 1.0f); outr = br; outi = bi; }
    if constexpr (MC > 0) { sc.run(p, br, bi, cr, ci, NST == 3 ? scale : 1.0f); outr = cr; outi = ci; }
    if constexpr (MD > 0) { sd.run(p, cr, ci, dr, di, NST == 4 ? scale : 1.0f); outr = dr; outi = di; }

    const long long i0 = q / D;  // index of the first output of this iteration
    if constexpr (!ARB) {
      if (i0 >= i_lo && i0 + NO <= i_hi && (NO % 2 == 0)) {
        float4* d4 = (float4*)(dst + (i0 & p.dst_mask));
#pragma unroll
        for (int b = 0; b < NO / 2; b++) d4[b] = make_float4(outr[2 * b], outi[2 * b], outr[2 * b + 1], outi[2 * b + 1]);
      } else {
#pragma unroll
        for (int b = 0; b < NO; b++)
          if (i0 + b >= i_lo && i0 + b < i_hi) dst[(i0 + b) & p.dst_mask] = make_float2(outr[b], outi[b]);
      }
    } else {
#pragma unroll
      for (int b = 0; b < NO; b++) { wr[13 + b] = outr[b]; wi[13 + b] = outi[b]; }
#pragma unroll
      for (int b = 0; b < NO; b++) {
        if (i0 + b >= i_lo && i0 + b < i_hi) {
          // A.5: emit outputs while phase < 2^24; newest sample is w[13 + b]
#pragma unroll 1
          while (phase < (1u << 24)) {
            const float4* row = (const float4*)(p.pfb + ((phase >> (24 - p.bits)) << 4));
            float4 h0 = __ldg(row), h1 = __ldg(row + 1), h2 = __ldg(row + 2), h3 = __ldg(row + 3);
            float h[14] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, h2.x, h2.y, h2.z, h2.w, h3.x, h3.y};
            float yr = 0.0f, yi = 0.0f;
#pragma unroll
            for (int k = 0; k < 14; k++) { yr = fmaf(h[k], wr[13 + b - k], yr); yi = fmaf(h[k], wi[13 + b - k], yi); }
            dst[j & p.dst_mask] = make_float2(yr, yi);
            j++;
            phase += p.step;
          }
          phase -= (1u << 24);
        }
      }
#pragma unroll
      for (int i = 0; i < 13; i++) { wr[i] = wr[NO + i]; wi[i] = wi[NO + i]; }
    }
  }
}

}  // namespace pmr
