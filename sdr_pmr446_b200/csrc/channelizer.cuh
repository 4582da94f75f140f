// channelize16_kernel: NCO mix-down + 16-channel polyphase analysis filter bank + 16-point DFT +
// NBFM discriminator.  Replaces the inner loop /root/reference/src/sdr_pmr446.c:804-823
// (nco_crcf_mix_down/step, firpfbch_crcf_analyzer_execute, transpose) and
// freqdem_demodulate_block (:881) for every channel.
//
// One warp walks a tile of 140 frames of one stream in five batches of 28 frames.
//  Phase A (two frames per step): lane l = (branch i = l & 15, half h = l >> 4); the 26-tap branch filter is split
//   in two 13-tap halves so that each lane keeps only a 14-slot window in registers (rotation period 14 frames =
//   7 steps, the unrolled loop body).  Half 1 runs on the same sample stream delayed by 13 frames, so both halves
//   execute the same code.  The commutator sends resampled sample 16f + 15 - i to branch i; the NCO phasor of a lane
//   only depends on the frame parity when 32 dtheta = 0 mod 2^32 (the reference's 17/32-cycle step), so mixing is one
//   complex multiply by a per-lane constant.  The halves swap one partial sum per step and the finished dot product
//   of frame k goes to shared memory, row k, column n = 15 - i (the DFT input index).
//  Phase B: lane = frame: reads its 16 branch outputs (row stride 17: conflict-free), runs the forward 16-point DFT in
//   registers (no shuffles) and writes channel-major rows [channel][1 + frame] (row stride 29).
//  Phase C: lane = (channel, 14 consecutive frames): discriminator arg(conj(y[f-1]) y[f]) with a polynomial atan2
//   (|err| < 3e-7 rad) and aligned two-sample stores into the channel's ring row; column 0 of each row carries the
//   last frame of the previous batch.
// The warp owns its shared-memory slice, so the phases are separated by __syncwarp() only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "audio_fft.cuh"   // dft16 / af_dig

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
struct ChanTaps {          // selector taps, only read by the TAPS instantiation
  float* mag_part;         // optional: [n_streams][tiles][16] sum of |y| over the owned frames of each tile (RSSI, :330-336)
  float2* edge;            // optional: [n_streams][16][2] channel sample of the first (f0) and last (f1 - 1) owned frame
};

constexpr int CH_TL = 136;      // owned frames per tile
constexpr int CH_BODIES = 10;   // 10 bodies x 14 frames = 140 computed frames, starting 4 before the tile

// atan2 by a degree-17 odd minimax polynomial on [0, 1] (Abramowitz & Stegun 4.4.49, |err| <= 2e-8 in
// exact arithmetic, < 3e-7 rad in float32); (0, 0) falls back to libm for the signed-zero cases.
__device__ __forceinline__ float fast_atan2f(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  // (+-0, +-0): the quotient is taken as 0, and the quadrant comes from the SIGN BITS below, which gives libm's
  // atan2(+-0, +0) = +-0 and atan2(+-0, -0) = +-pi without a branch
  const float a = __fdividef(mn, mx == 0.0f ? 1.0f : mx);
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
  if (__float_as_int(x) < 0) r = 3.14159265358979324f - r;
  return copysignf(r, y);
}

constexpr int CH_BATCH = 28;               // frames per batch: two rotations of the 14-slot window
constexpr int CH_NB = 5;                   // batches per tile (140 computed frames)
constexpr int CH_A_STRIDE = 17;            // float2 per frame row of the branch-output buffer (16 + 1 pad)
constexpr int CH_B_STRIDE = 29;            // float2 per channel row of the channel-output buffer (1 previous + 28)
constexpr int CH_SMEM_WARP = CH_BATCH * CH_A_STRIDE + 16 * CH_B_STRIDE;

template <bool TAPS> struct TapState { float macc = 0.0f; int k_first = -1, k_last = -1; };
template <> struct TapState<false> {};

// TAPS = the selector taps (mag_part / edge) are wanted: a separate instantiation, so that the throughput path carries
// none of their registers.
template <bool NCO_CONST, bool TAPS>
__global__ void __launch_bounds__(128, TAPS ? 4 : 5) channelize16_kernel(ChanParams p, ChanTaps tp) {
  __shared__ float2 ch_smem[4 * CH_SMEM_WARP];
  const int lane = threadIdx.x & 31, br = lane & 15, hsel = lane >> 4;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (long long)p.n_streams * p.tiles) return;   // warp-uniform
  float2* A = ch_smem + (threadIdx.x >> 5) * CH_SMEM_WARP;   // [frame of the batch][DFT input n]
  float2* B = A + CH_BATCH * CH_A_STRIDE;                     // [channel][previous frame, 28 frames]
  const int s = (int)(warp / p.tiles);
  const long long tile = p.tile0 + (warp % p.tiles);
  const long long fa = tile * CH_TL, fs = fa - 4;          // computed frames are fs + k, k in [0, 140)
  const float2* res = p.res + (long long)s * p.res_stride;

  float h[13];
#pragma unroll
  for (int n = 0; n < 13; n++) h[n] = __ldg(p.taps + br * 26 + 13 * hsel + n);

  // phase C of every batch: this lane owns channel br, frames 14 hsel .. 14 hsel + 13 of the batch
  float* drow = p.demod + ((long long)s * 16 + br) * p.demod_stride;
  float2* crow = p.chan ? p.chan + ((long long)s * 16 + br) * p.chan_ld : nullptr;

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
  // computed frame k is absolute frame 0 (stream start: r_prime = 0) when k == k_zero
  const int k_zero = (fs <= 0 && fs > -4096) ? (int)(-fs) : -1;
  if (lane < 16) B[lane * CH_B_STRIDE] = make_float2(0.0f, 0.0f);   // "previous frame" of the first batch (never owned)

  TapState<TAPS> ta;
  if constexpr (TAPS) {
    ta.k_first = (int)(p.f0 - fs > 4096 ? -1 : (p.f0 - fs < 0 ? -1 : p.f0 - fs));
    ta.k_last = (int)(p.f1 - 1 - fs > 4096 ? -1 : (p.f1 - 1 - fs < 0 ? -1 : p.f1 - 1 - fs));
  }
  float2 n0 = fetch(0), n1 = fetch(1);
#pragma unroll 1
  for (int batch = 0; batch < CH_NB; batch++) {
    // ---- phase A: mix, branch filters; dot products of frame kk go to A[kk][15 - branch] ----------------------------
#pragma unroll 1
    for (int hb = 0; hb < 2; hb++) {
      const int kb = CH_BATCH * batch + 14 * hb;
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
        A[(14 * hb + 2 * j + hsel) * CH_A_STRIDE + 15 - br] = make_float2(ar, ai);
      }
    }
    __syncwarp();
    // ---- phase B: lane = frame of the batch: forward 16-point DFT in registers, transposed to B[channel][1 + frame] --
    if (lane < CH_BATCH) {
      float2 v[16];
#pragma unroll
      for (int n = 0; n < 16; n++) v[n] = A[lane * CH_A_STRIDE + n];
      dft16<false>(v);
#pragma unroll
      for (int c = 0; c < 16; c++) B[c * CH_B_STRIDE + 1 + lane] = v[af_dig(c)];
    }
    __syncwarp();
    // ---- phase C: lane = (channel br, frames 14 hsel ..): discriminator (A.9) arg(conj(prev) y) * ref with separate
    // mul/add like the C reference, two-sample stores --------------------------------------------------------------------
    {
      const int k0 = CH_BATCH * batch + 14 * hsel;
      const float2* brow = B + br * CH_B_STRIDE + 14 * hsel;
      float2 prev = brow[0];
      const bool all = k0 >= own_lo && k0 + 14 <= own_hi;
      // (fs + k0) is even: a pair of frames is an aligned float2 of the ring and never straddles its end
#pragma unroll 1
      for (int i = 0; i < 14; i += 2) {
        const float2 y0 = brow[1 + i], y1 = brow[2 + i];
        if (k0 + i == k_zero) prev = make_float2(0.0f, 0.0f);
        const float re0 = __fadd_rn(__fmul_rn(prev.x, y0.x), __fmul_rn(prev.y, y0.y));
        const float im0 = __fsub_rn(__fmul_rn(prev.x, y0.y), __fmul_rn(prev.y, y0.x));
        float2 p1 = y0;
        if (k0 + i + 1 == k_zero) p1 = make_float2(0.0f, 0.0f);
        const float re1 = __fadd_rn(__fmul_rn(p1.x, y1.x), __fmul_rn(p1.y, y1.y));
        const float im1 = __fsub_rn(__fmul_rn(p1.x, y1.y), __fmul_rn(p1.y, y1.x));
        const float2 dm = make_float2(fast_atan2f(im0, re0) * p.ref, fast_atan2f(im1, re1) * p.ref);
        const int k = k0 + i;
        if constexpr (TAPS) {
          if (tp.mag_part) {
            if (k >= own_lo && k < own_hi) ta.macc += sqrtf(fmaf(y0.x, y0.x, y0.y * y0.y));
            if (k + 1 >= own_lo && k + 1 < own_hi) ta.macc += sqrtf(fmaf(y1.x, y1.x, y1.y * y1.y));
          }
          if (tp.edge) {
            float2* e = tp.edge + ((long long)s * 16 + br) * 2;
            if (k == ta.k_first) e[0] = y0;
            if (k + 1 == ta.k_first) e[0] = y1;
            if (k == ta.k_last) e[1] = y0;
            if (k + 1 == ta.k_last) e[1] = y1;
          }
        }
        float* d = drow + ((fs32 + (unsigned)k) & dmask);
        if (all) {
          *(float2*)d = dm;
          if (crow) { crow[crel + k] = y0; crow[crel + k + 1] = y1; }
        } else {
          const bool a0 = k >= own_lo && k < own_hi, a1 = k + 1 >= own_lo && k + 1 < own_hi;
          if (a0) d[0] = dm.x;
          if (a1) d[1] = dm.y;
          if (crow) {
            if (a0) crow[crel + k] = y0;
            if (a1) crow[crel + k + 1] = y1;
          }
        }
        prev = y1;
      }
      __syncwarp();
      if (hsel) B[br * CH_B_STRIDE] = prev;      // frame 27 of this batch is the next batch's previous frame
    }
    __syncwarp();
  }
  if constexpr (TAPS) {
    if (tp.mag_part) {
      const float m = ta.macc + __shfl_xor_sync(0xffffffffu, ta.macc, 16);
      if (hsel == 0) tp.mag_part[((long long)s * p.tiles + (warp % p.tiles)) * 16 + br] = m;
    }
  }
}

// rssi[s][c] = 20 log10(sum over the tiles of mag_part / ns): average_power() of the reference for every channel
static __global__ void rssi_finalize_kernel(const float* mag_part, int tiles, int rows, long long ns, float* rssi) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;   // r = s * 16 + c
  if (r >= rows) return;
  const float* m = mag_part + (long long)(r >> 4) * tiles * 16 + (r & 15);
  float acc = 0.0f;
  for (int t = 0; t < tiles; t++) acc += m[(long long)t * 16];
  rssi[r] = 20.0f * log10f(acc / (float)ns);
}

}  // namespace pmr
