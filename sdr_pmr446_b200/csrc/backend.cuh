// Back-end kernels of the PMR446 chain.
//
//  channelize16_kernel : NCO mix-down + 16-channel polyphase analysis filter bank + 16-point DFT
//                        + NBFM discriminator.  Replaces the inner loop
//                        /root/reference/src/sdr_pmr446.c:804-823 (nco_crcf_mix_down/step,
//                        firpfbch_crcf_analyzer_execute, transpose) and freqdem_demodulate_block
//                        (:881) for every channel.
//  audio_kernel        : 377-tap CTCSS-removal high-pass FIR, complementary low-pass branch,
//                        audio gain, 1-pole de-emphasis, optional 103-tap low-pass, s16 conversion.
//                        Replaces :882-902 (firfilt_rrrf_execute_block, wdelayf, iirfilt_rrrf)
//                        (Appendix A.1, A.10, A.11) and the s16 cast of src/dsd_in.c:172-175.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmr {

// ------------------------------------------------------------------------------------------
// Channelizer.  One half-warp (16 lanes) walks a tile of frames of one stream; lane i is
// polyphase branch i: it keeps its 26-sample window in registers (28 slots so the rotation has
// period 28 = 7 float4 store groups), takes one mixed sample per frame (the commutator sends
// sample 16f + 15 - i to branch i), does the 26-tap dot product and joins a 4-stage
// decimation-in-frequency FFT over the 16 lanes with xor-shuffles.  After the FFT the lane with
// DFT input index n = 15 - i holds bin bitrev4(n), i.e. PMR channel c = bitrev4(15 - i); it runs
// the discriminator for that channel on consecutive frames and writes 4 frames per float4.
struct ChanParams {
  const float2* res;       // resampler output ring [n_streams][res_stride]
  long long res_stride, res_mask;
  long long r1;            // resampler outputs available: [.., r1)
  int n_streams;
  int tiles;               // tiles per stream
  long long tile0;         // index of the first tile (tile k covers frames [k*TL, (k+1)*TL))
  long long f0, f1;        // owned frames [f0, f1)
  unsigned dtheta;         // NCO phase increment per sample (A.7)
  float ref;               // 1 / (2 pi kf)
  const float* taps;       // [16][26] branch taps, newest first
  float* demod;            // ring [n_streams*16][demod_stride]
  long long demod_stride, demod_mask;
  float2* chan;            // optional: [n_streams][16][chan_ld], column = f - f0
  long long chan_ld;
};

constexpr int CH_TL = 136;      // owned frames per tile (28*5 - 4)
constexpr int CH_BLOCKS28 = 5;

__device__ __forceinline__ float fast_atan2f(float y, float x) { return atan2f(y, x); }

template <bool NCO_LUT>
__global__ void __launch_bounds__(128) channelize16_kernel(ChanParams p) {
  __shared__ float2 lut[32];
  if (NCO_LUT) {
    if (threadIdx.x < 32) {
      float sn, cs;
      sincospif((float)threadIdx.x * (1.0f / 16.0f), &sn, &cs);
      lut[threadIdx.x] = make_float2(cs, sn);
    }
    __syncthreads();
  }
  const int lane16 = threadIdx.x & 15;
  long long grp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const bool active = grp < (long long)p.n_streams * p.tiles;
  if (!active) grp = 0;  // keep the lanes alive for the shuffles; stores are masked below
  const int s = (int)(grp / p.tiles);
  const long long tile = p.tile0 + (grp % p.tiles);
  const long long fa = tile * CH_TL;
  const float2* res = p.res + (long long)s * p.res_stride;

  float h[26];
#pragma unroll
  for (int n = 0; n < 26; n++) h[n] = __ldg(p.taps + lane16 * 26 + n);

  // FFT constants of this lane: DFT input index n = 15 - lane16
  const int nidx = 15 - lane16;
  float sg[4], twr[3], twi[3];
#pragma unroll
  for (int st = 0; st < 4; st++) {
    const int hh = 8 >> st;
    const bool hi = (nidx & hh) != 0;
    sg[st] = hi ? -1.0f : 1.0f;
    if (st < 3) {
      float sn = 0.0f, cs = 1.0f;
      if (hi) sincospif(-(float)(nidx & (hh - 1)) / (float)hh, &sn, &cs);
      twr[st] = cs;
      twi[st] = sn;
    }
  }
  const int c = ((nidx & 1) << 3) | ((nidx & 2) << 1) | ((nidx & 4) >> 1) | ((nidx & 8) >> 3);
  float* drow = p.demod + ((long long)s * 16 + c) * p.demod_stride;
  float2* crow = p.chan ? p.chan + ((long long)s * 16 + c) * p.chan_ld : nullptr;

  auto fetch = [&](long long f, float& xr, float& xi) {
    const long long j = 16 * f + 15 - lane16;
    float2 v = make_float2(0.0f, 0.0f);
    if (j >= 0 && j < p.r1) v = res[j & p.res_mask];
    float cs, sn;
    const unsigned th = (unsigned)j * p.dtheta;
    if (NCO_LUT) {
      float2 e = lut[th >> 27];
      cs = e.x; sn = e.y;
    } else {
      sincospif((float)(int)th * (1.0f / 2147483648.0f), &sn, &cs);
    }
    xr = fmaf(v.x, cs, v.y * sn);    // v * conj(e^{j theta})
    xi = fmaf(v.y, cs, -v.x * sn);
  };

  float wr[28], wi[28];
  const long long fs = fa - 4;
  wr[0] = wi[0] = wr[1] = wi[1] = wr[2] = wi[2] = 0.0f;
#pragma unroll
  for (int n = 1; n <= 25; n++) fetch(fs - n, wr[28 - n], wi[28 - n]);

  float pr = 0.0f, pi = 0.0f;  // previous channel sample (discriminator state r_prime)
  float dm[4];
  float2 ch[4];
#pragma unroll 1
  for (int blk = 0; blk < CH_BLOCKS28; blk++) {
    const long long fb = fs + 28 * blk;
#pragma unroll
    for (int ff = 0; ff < 28; ff++) {
      const long long f = fb + ff;
      fetch(f, wr[ff], wi[ff]);
      float ar = 0.0f, ai = 0.0f;
#pragma unroll
      for (int n = 0; n < 26; n++) {
        ar = fmaf(h[n], wr[(ff - n + 28) % 28], ar);
        ai = fmaf(h[n], wi[(ff - n + 28) % 28], ai);
      }
      // 16-point DIF FFT across the half-warp
#pragma unroll
      for (int st = 0; st < 4; st++) {
        const int hh = 8 >> st;
        float br = __shfl_xor_sync(0xffffffffu, ar, hh);
        float bi = __shfl_xor_sync(0xffffffffu, ai, hh);
        float tr = fmaf(sg[st], ar, br), ti = fmaf(sg[st], ai, bi);
        if (st < 3) {
          ar = fmaf(tr, twr[st], -ti * twi[st]);
          ai = fmaf(tr, twi[st], ti * twr[st]);
        } else {
          ar = tr; ai = ti;
        }
      }
      // discriminator (A.9): arg(conj(prev) * y) * ref
      // (separate mul/add like the C reference, so signed zeros at stream start behave the same)
      if (f == 0) { pr = 0.0f; pi = 0.0f; }
      const float re = __fadd_rn(__fmul_rn(pr, ar), __fmul_rn(pi, ai));
      const float im = __fsub_rn(__fmul_rn(pr, ai), __fmul_rn(pi, ar));
      dm[ff & 3] = fast_atan2f(im, re) * p.ref;
      ch[ff & 3] = make_float2(ar, ai);
      pr = ar; pi = ai;
      if ((ff & 3) == 3 && active) {
        const long long g0 = f - 3;
        if (g0 >= p.f0 && f < p.f1 && g0 >= fa && f < fa + CH_TL) {
          *(float4*)(drow + (g0 & p.demod_mask)) = make_float4(dm[0], dm[1], dm[2], dm[3]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const long long fk = g0 + k;
            if (fk >= p.f0 && fk < p.f1 && fk >= fa && fk < fa + CH_TL) drow[fk & p.demod_mask] = dm[k];
          }
        }
        if (crow) {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const long long fk = g0 + k;
            if (fk >= p.f0 && fk < p.f1 && fk >= fa && fk < fa + CH_TL) crow[fk - p.f0] = ch[k];
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Audio chain.  One block = one (stream, channel) row x one time tile.  Thread t owns 16
// consecutive output samples; the tile's input (with a halo of taps-1 samples) sits in shared
// memory with one pad word every 16 so that the per-thread sliding windows (lane stride 17
// words) are bank-conflict free.  The FIR runs 16 taps x 16 outputs per step from registers:
// 256 FFMA per 16 LDS (samples) + 4 LDS.128 (broadcast taps).
struct AudioParams {
  const float* demod;      // ring [rows][demod_stride]
  long long demod_stride, demod_mask;
  int rows;                // n_streams * 16
  int tiles;               // tiles per row
  long long tile0;         // tile k covers outputs [k*TT, (k+1)*TT), TT = AU_OWN
  long long f0, f1;        // owned samples
  const float* hp_taps;    // [HP_PAD] zero-padded to a multiple of 16
  int hp_chunks;           // HP_PAD / 16
  int hp_delay;            // (hp_len - 1) / 2
  const float* lp_taps;    // optional second FIR (103 taps padded); nullptr = off
  int lp_chunks;
  float gain;
  float de_b0, de_b1, de_a1;
  float* audio;            // optional [rows][out_ld] float audio, column = f - f0
  short* pcm;              // optional [rows][out_ld]
  float* lpcomp;           // optional [rows][out_ld]
  long long out_ld;
};

constexpr int AU_THREADS = 128;
constexpr int AU_SPAN = AU_THREADS * 16;   // outputs computed per tile (incl. lead-in)
constexpr int AU_LEAD = 128;               // lead-in outputs (de-emphasis warm-up + LP halo), multiple of 16
constexpr int AU_OWN = AU_SPAN - AU_LEAD;  // owned outputs per tile
constexpr int AU_MAXHALO = 384;            // >= padded HP length

__device__ __forceinline__ int padidx(int i) { return i + (i >> 4); }

// acc[r] += sum_k taps[k] * x[base + r - k], r < 16, k < 16*chunks, for this thread.
// xs is the padded tile; `top` = padded-layout-free index of x[base] (multiple of 16).
__device__ __forceinline__ void fir16(const float* __restrict__ xs, const float* __restrict__ ts, int top, int chunks, float* acc) {
  // window registers: lo = x[top - 16kb - 15 .. top - 16kb], hi = previous lo
  float wa[16], wb[16];
  // first chunk needs x[top .. top + 15] as the "hi" part (r - kk > 0)
  {
    const float* q = xs + padidx(top);
#pragma unroll
    for (int i = 0; i < 16; i++) wb[i] = q[i];   // top is a multiple of 16: no pad inside
  }
  auto step = [&](float* lo, const float* hi, int kb) {
    // lo[i] = x[top - 16(kb+1) + i], i < 16 ; hi[i] = x[top - 16 kb + i]
    const float* q = xs + padidx(top - 16 * (kb + 1));
#pragma unroll
    for (int i = 0; i < 16; i++) lo[i] = q[i];
    const float4* t4 = (const float4*)(ts + 16 * kb);
    float4 t0 = t4[0], t1 = t4[1], t2 = t4[2], t3 = t4[3];
    const float tk[16] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w, t3.x, t3.y, t3.z, t3.w};
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
#pragma unroll
      for (int r = 0; r < 16; r++) {
        const int d = r - kk;  // x index relative to top - 16 kb
        const float xv = d >= 0 ? hi[d] : lo[16 + d];
        acc[r] = fmaf(tk[kk], xv, acc[r]);
      }
    }
  };
  int kb = 0;
#pragma unroll 1
  for (; kb + 1 < chunks; kb += 2) {
    step(wa, wb, kb);
    step(wb, wa, kb + 1);
  }
  if (kb < chunks) step(wa, wb, kb);
}

__global__ void __launch_bounds__(AU_THREADS) audio_kernel(AudioParams p) {
  extern __shared__ float smem[];
  // layout: xs[padidx(AU_MAXHALO + AU_SPAN + 16)] input tile; ys[...] second buffer (gain*hp, then de-emph)
  constexpr int XN = AU_MAXHALO + AU_SPAN + 16;
  float* xs = smem;
  float* ys = smem + padidx(XN) + 1;
  float* ts = ys + padidx(AU_LEAD + AU_SPAN + 16) + 1;   // taps (16-byte aligned below)
  ts = (float*)(((uintptr_t)ts + 15) & ~(uintptr_t)15);

  const int row = blockIdx.x / p.tiles;
  const long long tile = p.tile0 + blockIdx.x % p.tiles;
  const long long o0 = tile * AU_OWN - AU_LEAD;   // absolute index of computed output 0 (multiple of 16)
  const float* drow = p.demod + (long long)row * p.demod_stride;
  const int t = threadIdx.x;

  // load x[o0 - AU_MAXHALO .. o0 + AU_SPAN) ; negative absolute indices are zero (stream start)
  for (int i = t; i < AU_MAXHALO + AU_SPAN; i += AU_THREADS) {
    const long long n = o0 - AU_MAXHALO + i;
    xs[padidx(i)] = (n >= 0 && n < p.f1) ? drow[n & p.demod_mask] : 0.0f;
  }
  const int hp_n = p.hp_chunks * 16, lp_n = p.lp_taps ? p.lp_chunks * 16 : 0;
  for (int i = t; i < hp_n; i += AU_THREADS) ts[i] = p.hp_taps[i];
  for (int i = t; i < lp_n; i += AU_THREADS) ts[hp_n + i] = p.lp_taps[i];
  __syncthreads();

  // high-pass FIR: outputs o0 + 16 t + r
  float acc[16];
#pragma unroll
  for (int r = 0; r < 16; r++) acc[r] = 0.0f;
  const int top = AU_MAXHALO + 16 * t;
  fir16(xs, ts, top, p.hp_chunks, acc);

  const long long obase = o0 + 16 * t;
  const bool own_thread = (16 * t >= AU_LEAD);
  // complementary low-pass branch (A.11): delayed input minus high-pass output
  if (p.lpcomp && own_thread) {
    float* lrow = p.lpcomp + (long long)row * p.out_ld;
#pragma unroll
    for (int r = 0; r < 16; r++) {
      const long long f = obase + r;
      if (f >= p.f0 && f < p.f1) lrow[f - p.f0] = xs[padidx(top + r - p.hp_delay)] - acc[r];
    }
  }
  // gain, then de-emphasis (A.1, Direct Form II): v = x - a1 v1 ; y = b0 v + b1 v1
#pragma unroll
  for (int r = 0; r < 16; r++) {
    acc[r] *= p.gain;
    ys[padidx(AU_LEAD + 16 * t + r)] = acc[r];
  }
  __syncthreads();
  {
    // pole magnitude is 0.0146: eight samples of warm-up reproduce the running state to < 1e-14
    float v1 = 0.0f;
#pragma unroll
    for (int k = 8; k >= 1; k--) {
      const int idx = AU_LEAD + 16 * t - k;
      const float xv = (idx >= AU_LEAD) ? ys[padidx(idx)] : 0.0f;
      v1 = fmaf(-p.de_a1, v1, xv);
    }
#pragma unroll
    for (int r = 0; r < 16; r++) {
      const float v0 = fmaf(-p.de_a1, v1, acc[r]);
      acc[r] = fmaf(p.de_b0, v0, p.de_b1 * v1);
      v1 = v0;
    }
  }
  if (p.lp_taps) {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) ys[padidx(AU_LEAD + 16 * t + r)] = acc[r];
    if (t < AU_LEAD / 16) {
#pragma unroll
      for (int r = 0; r < 16; r++) ys[padidx(16 * t + r)] = 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) acc[r] = 0.0f;
    fir16(ys, ts + hp_n, AU_LEAD + 16 * t, p.lp_chunks, acc);
  }
  // stage through shared memory for coalesced, arbitrarily aligned stores
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 16; r++) xs[padidx(16 * t + r)] = acc[r];
  __syncthreads();
  float* arow = p.audio ? p.audio + (long long)row * p.out_ld : nullptr;
  short* prow = p.pcm ? p.pcm + (long long)row * p.out_ld : nullptr;
  const long long own_lo = o0 + AU_LEAD, own_hi = o0 + AU_SPAN;
  for (int i = AU_LEAD + t; i < AU_SPAN; i += AU_THREADS) {
    const long long f = o0 + i;
    if (f >= p.f0 && f < p.f1 && f >= own_lo && f < own_hi) {
      const float v = xs[padidx(i)];
      if (arow) arow[f - p.f0] = v;
      if (prow) prow[f - p.f0] = (short)__float2int_rz(v * 32767.0f);
    }
  }
}

}  // namespace pmr
