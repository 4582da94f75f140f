// channelize16_kernel: NCO mix-down + 16-channel polyphase analysis filter bank + 16-point DFT +
// NBFM discriminator.  Replaces the inner loop /root/reference/src/sdr_pmr446.c:804-823
// (nco_crcf_mix_down/step, firpfbch_crcf_analyzer_execute, transpose) and
// freqdem_demodulate_block (:881) for every channel.
//
// One warp walks a tile of frames of one stream, two frames per step.  Lane l = (branch i = l & 15,
// half h = l >> 4): the 26-tap branch filter is split in two 13-tap halves so that each lane keeps
// only a 14-slot window in registers (rotation period 14 frames = 7 steps, the unrolled loop body).
// Half 1 simply runs on the same sample stream delayed by 13 frames, so both halves execute the
// same code.  The commutator sends resampled sample 16f + 15 - i to branch i; the NCO phasor of a
// lane only depends on the frame parity when 32 dtheta = 0 mod 2^32 (the reference's 17/32-cycle
// step), so mixing is one complex multiply by a per-lane constant.  Per step the halves swap one
// partial sum (half 0 finishes the even frame, half 1 the odd one), each 16-lane half runs a
// 4-stage decimation-in-frequency FFT with xor-shuffles (the lane holding DFT input n = 15 - i ends
// with bin bitrev4(n), i.e. PMR channel c), and the discriminator arg(conj(y[f-1]) y[f]) runs with
// the previous frame's value obtained from the other half.  A polynomial atan2 (|err| < 3e-7 rad)
// replaces libm's.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmr {

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

constexpr int CH_TL = 136;      // owned frames per tile
constexpr int CH_BODIES = 10;   // 10 bodies x 14 frames = 140 computed frames, starting 4 before the tile

// atan2 by a degree-17 odd minimax polynomial on [0, 1] (Abramowitz & Stegun 4.4.49, |err| <= 2e-8 in
// exact arithmetic, < 3e-7 rad in float32); (0, 0) falls back to libm for the signed-zero cases.
__device__ __forceinline__ float fast_atan2f(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  if (mx == 0.0f) return atan2f(y, x);
  const float a = __fdividef(mn, mx);
  const float s = a * a;
  float r = 0.0028662257f;
  r = fmaf(r, s, -0.0161657367f);
  r = fmaf(r, s, 0.0429096138f);
  r = fmaf(r, s, -0.0752896400f);
  r = fmaf(r, s, 0.1065626393f);
  r = fmaf(r, s, -0.1420889944f);
  r = fmaf(r, s, 0.1999355085f);
  r = fmaf(r, s, -0.3333314528f);
  r = fmaf(r, s, 1.0f);
  r *= a;
  if (ay > ax) r = 1.57079632679489662f - r;
  if (x < 0.0f) r = 3.14159265358979324f - r;
  return copysignf(r, y);
}

template <bool NCO_CONST>
__global__ void __launch_bounds__(128) channelize16_kernel(ChanParams p) {
  const int lane = threadIdx.x & 31, br = lane & 15, hsel = lane >> 4;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (long long)p.n_streams * p.tiles) return;   // warp-uniform
  const int s = (int)(warp / p.tiles);
  const long long tile = p.tile0 + (warp % p.tiles);
  const long long fa = tile * CH_TL, fs = fa - 4;          // computed frames are fs + k, k in [0, 140)
  const float2* res = p.res + (long long)s * p.res_stride;

  float h[13];
#pragma unroll
  for (int n = 0; n < 13; n++) h[n] = __ldg(p.taps + br * 26 + 13 * hsel + n);

  // FFT constants of this lane: DFT input index n = 15 - branch
  const int nidx = 15 - br;
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

  // this lane's sample for (delayed) frame k: j = jb + 16 k; valid iff 0 <= j < r1
  const long long jb = 16 * (fs - 13 * hsel) + 15 - br;
  const unsigned jb32 = (unsigned)jb, rmask = (unsigned)p.res_mask;
  int k_lo = -(1 << 20), k_hi = 1 << 20;
  if (jb < 0) k_lo = (int)((-jb + 15) / 16);
  {
    const long long t = p.r1 - jb;   // j < r1  <=>  16 k < t
    if (t <= 0) k_hi = -(1 << 20);
    else if (t < (1ll << 24)) k_hi = (int)((t + 15) / 16);
  }
  // NCO phasors (A.7): theta_j = j * dtheta mod 2^32; constant per frame parity when 32 dtheta = 0
  float pc[2], ps[2];
#pragma unroll
  for (int par = 0; par < 2; par++) {
    const unsigned th = (jb32 + 16u * (unsigned)par) * p.dtheta;
    sincospif((float)(int)th * (1.0f / 2147483648.0f), &ps[par], &pc[par]);
  }
  auto fetch = [&](int k) -> float2 {   // branch-free: out-of-range frames read slot 0 of the ring and are zeroed
    const bool ok = k >= k_lo && k < k_hi;
    const unsigned idx = ok ? ((jb32 + 16u * (unsigned)k) & rmask) : 0u;
    float2 v = res[idx];
    if (!ok) v = make_float2(0.0f, 0.0f);
    return v;
  };
  auto mix = [&](float2 v, int k, int par, float& xr, float& xi) {   // par = k & 1, passed as a literal
    float cs, sn;
    if (NCO_CONST) {
      cs = pc[par]; sn = ps[par];
    } else {
      const unsigned th = (jb32 + 16u * (unsigned)k) * p.dtheta;
      sincospif((float)(int)th * (1.0f / 2147483648.0f), &sn, &cs);
    }
    xr = fmaf(v.x, cs, v.y * sn);    // v * conj(e^{j theta})
    xi = fmaf(v.y, cs, -v.x * sn);
  };

  // window slot of frame k is k mod 14; preload frames -12..-1 into slots 2..13
  float wr[14], wi[14];
  wr[0] = wi[0] = wr[1] = wi[1] = 0.0f;
#pragma unroll
  for (int n = 1; n <= 12; n++) mix(fetch(-n), -n, n & 1, wr[14 - n], wi[14 - n]);

  // ownership of computed frame k (relative to fs), 32-bit
  long long lo64 = (fa > p.f0 ? fa : p.f0) - fs, hi64 = ((fa + CH_TL) < p.f1 ? (fa + CH_TL) : p.f1) - fs;
  const int own_lo = (int)(lo64 < 0 ? 0 : (lo64 > 4096 ? 4096 : lo64)), own_hi = (int)(hi64 < 0 ? 0 : (hi64 > 4096 ? 4096 : hi64));
  const unsigned fs32 = (unsigned)fs, dmask = (unsigned)p.demod_mask;
  const int crel = (int)(fs - p.f0 < -(1ll << 30) ? -(1 << 30) : (fs - p.f0 > (1ll << 30) ? (1 << 30) : fs - p.f0));
  // computed frame k of this lane is absolute frame 0 (stream start: r_prime = 0) when k == k_zero
  const long long kz64 = -fs - hsel;
  const int k_zero = (int)(kz64 < -1 ? -1 : (kz64 > 4096 ? -1 : kz64));

  float sav_r = 0.0f, sav_i = 0.0f;   // half 0: y of the previous odd frame (from half 1)
  float2 n0 = fetch(0), n1 = fetch(1);
#pragma unroll 1
  for (int body = 0; body < CH_BODIES; body++) {
    const int kb = 14 * body;
#pragma unroll
    for (int j = 0; j < 7; j++) {
      const int k = kb + 2 * j;            // frames k (slot 2j) and k + 1 (slot 2j + 1)
      mix(n0, k, 0, wr[2 * j], wi[2 * j]);
      mix(n1, k + 1, 1, wr[2 * j + 1], wi[2 * j + 1]);
      n0 = fetch(k + 2);                   // one step ahead
      n1 = fetch(k + 3);
      float d0r = 0.0f, d0i = 0.0f, d1r = 0.0f, d1i = 0.0f;
#pragma unroll
      for (int n = 0; n < 13; n++) {
        d0r = fmaf(h[n], wr[(2 * j - n + 14) % 14], d0r);
        d0i = fmaf(h[n], wi[(2 * j - n + 14) % 14], d0i);
        d1r = fmaf(h[n], wr[(2 * j + 1 - n + 14) % 14], d1r);
        d1i = fmaf(h[n], wi[(2 * j + 1 - n + 14) % 14], d1i);
      }
      // half 0 completes frame k, half 1 frame k + 1: swap the other frame's partial sum
      float ar = hsel ? d1r : d0r, ai = hsel ? d1i : d0i;
      ar += __shfl_xor_sync(0xffffffffu, hsel ? d0r : d1r, 16);
      ai += __shfl_xor_sync(0xffffffffu, hsel ? d0i : d1i, 16);
      // 16-point DIF FFT within each half
#pragma unroll
      for (int st = 0; st < 4; st++) {
        const int hh = 8 >> st;
        const float br_ = __shfl_xor_sync(0xffffffffu, ar, hh);
        const float bi_ = __shfl_xor_sync(0xffffffffu, ai, hh);
        const float tr = fmaf(sg[st], ar, br_), ti = fmaf(sg[st], ai, bi_);
        if (st < 3) {
          ar = fmaf(tr, twr[st], -ti * twi[st]);
          ai = fmaf(tr, twi[st], ti * twr[st]);
        } else {
          ar = tr; ai = ti;
        }
      }
      // discriminator (A.9): arg(conj(prev) * y) * ref with separate mul/add like the C reference
      const float exr = __shfl_xor_sync(0xffffffffu, ar, 16), exi = __shfl_xor_sync(0xffffffffu, ai, 16);
      float pr = hsel ? exr : sav_r, pi = hsel ? exi : sav_i;
      sav_r = exr; sav_i = exi;
      if (k == k_zero) { pr = 0.0f; pi = 0.0f; }
      const float re = __fadd_rn(__fmul_rn(pr, ar), __fmul_rn(pi, ai));
      const float im = __fsub_rn(__fmul_rn(pr, ai), __fmul_rn(pi, ar));
      const float dm = fast_atan2f(im, re) * p.ref;
      const float dm_o = __shfl_xor_sync(0xffffffffu, dm, 16);
      // half 0 stores frames k, k + 1 (own value and the other half's); (fs + k) is even, so the pair
      // is an aligned float2 unless the ownership boundary splits it
      {
        const bool a0 = (hsel == 0) && k >= own_lo && k < own_hi, a1 = (hsel == 0) && k + 1 >= own_lo && k + 1 < own_hi;
        float* d = drow + ((fs32 + (unsigned)k) & dmask);
        if (a0) d[0] = dm;
        if (a1) d[1] = dm_o;
        if (crow) {
          if (a0) crow[crel + k] = make_float2(ar, ai);
          if (a1) crow[crel + k + 1] = make_float2(exr, exi);
        }
      }
    }
  }
}

}  // namespace pmr
