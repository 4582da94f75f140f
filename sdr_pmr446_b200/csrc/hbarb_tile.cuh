// hbarb_tile_kernel: last half-band decimator + arbitrary resampler of msresamp_crcf for plans whose
// resampler phase is periodic with a short period (step * P = Q * 2^24; the 2.4 Msps plan has rate 2/3:
// P = 2 outputs per Q = 3 inputs).  Same arithmetic as cascade_kernel<SRC_RING, DC_NONE, 8, M, 0, 0, 0, true>
// (/root/reference/src/sdr_pmr446.c:796 via liquid's msresamp2 + resamp), other mapping:
//
//   one block = one stream x one tile of 768 resampler outputs (absolute grid, so chunk boundaries only move the
//   ownership window [j0, j1)).  The 600 kHz ring samples of the tile (+ filter history) are staged in shared
//   memory split in even / odd phases, with the zero-input-response correction of the DC blocker applied on the
//   way in; every thread then computes 9 consecutive half-band outputs from a register window (28 even samples,
//   20 taps as constant-bank operands) into shared memory -- reusing the even-phase buffer -- and 6 consecutive
//   resampler outputs from 21 of those; the two phase rows of the filter bank are kernel parameters.  With 6 outputs
//   per thread the lane stride is 9 elements (odd): conflict-free 64-bit accesses without padding, 64 registers,
//   19 KB of shared memory, 8 blocks per SM.  (8 outputs per thread needed one pad per 12 elements, whose jumps cost
//   a bank conflict on every staging store: 1.44 vs 1.39 ms; 10 per thread: 126 registers, 1.49 ms.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "frontend.cuh"

namespace pmr {

constexpr int HT_THREADS = 128;
constexpr int HT_R = 6;                  // resampler outputs per thread
constexpr int HT_TJ = HT_THREADS * HT_R;   // resampler outputs per tile
constexpr int HT_HH = 16;                  // half-band outputs recomputed before the tile (>= 13 taps of history)

struct HbArbParams {
  const float2* ring;       // producer ring [n_streams][ring_stride]
  long long ring_stride, ring_mask;
  long long n1;             // ring samples available: [.., n1)
  Correction corr;
  int n_streams, tiles;     // tiles per stream in this launch
  long long tile0;          // absolute index of the first tile
  long long j0, j1;         // owned resampler outputs
  float scale;              // 2^-stages, applied to the half-band output like msresamp does
  float hb[20];             // half-band taps h[j], j < 2M, newest first
  float arb[2][14];         // filter-bank rows of the P phases
  float2* dst;              // resampled ring
  long long dst_stride, dst_mask;
};

template <int M, int P, int Q>
__global__ void __launch_bounds__(HT_THREADS, 8) hbarb_tile_kernel(HbArbParams p) {
  static_assert(P == 2 && HT_R % P == 0, "phase rows are passed for P = 2");
  constexpr int RH = HT_R / P * Q;          // half-band outputs per thread (12)
  constexpr int TH = RH * HT_THREADS;       // per tile
  constexpr int W = 2 * M - 1;              // previous even samples a half-band output needs
  constexpr int NPAIR = TH + HT_HH + W;     // (even, odd) pairs staged per tile
  // lane stride in float2 units must be odd for conflict-free 64-bit accesses: RH itself when odd, else one pad per RH
  constexpr int PADK = (RH % 2 == 0) ? 1 : 0;
  auto pad = [](int i) { return i + PADK * (i / RH); };
  extern __shared__ float2 ht_smem[];
  float2* ev = ht_smem;                      // even-phase ring samples
  float2* od = ev + pad(NPAIR) + 1;          // odd-phase
  float2* hbo = ev;                          // half-band outputs overwrite the even-phase tile (after a barrier)

  const int t = threadIdx.x;
  const int s = blockIdx.x / p.tiles;
  const long long tile = p.tile0 + blockIdx.x % p.tiles;
  const long long jt = tile * HT_TJ;          // first resampler output of the tile
  const long long A = tile * TH;              // its first half-band output (jt / P * Q)
  const long long pair0 = A - HT_HH - W;      // absolute pair index of ev[0] / od[0]

  // ---- stage the ring tile: pair q = samples 2 (pair0 + q), 2 (pair0 + q) + 1 ----------------------------------
  {
    const float2* ring = p.ring + (long long)s * p.ring_stride;
    const Correction& c = p.corr;
    const float2* vrow = c.v_seg ? c.v_seg + (long long)s * c.nseg : nullptr;
    // block-uniform 32-bit bounds, relative to the tile's first sample nb = 2 pair0 (may be negative at stream start)
    const long long nb = 2 * pair0;
    auto clampi = [](long long v) { return (int)(v < -(1 << 29) ? -(1 << 29) : (v > (1 << 29) ? (1 << 29) : v)); };
    const int q_lo = nb < 0 ? (int)((-nb + 1) / 2) : 0;          // first pair with n >= 0
    const int avail = clampi(p.n1 - nb);                         // samples of the tile that exist
    const int q_full = avail >> 1;                               // pairs [q_lo, q_full) exist completely
    const int from_s = clampi(c.from - nb);                      // samples >= from_s still need the correction
    const int seg_rel = clampi(nb - c.seg0_out);                 // tile start relative to the producer's segment grid
    const unsigned nb32 = (unsigned)nb, rmask = (unsigned)p.ring_mask, smask = (1u << c.seg_shift) - 1u;
    // all loads of the thread are issued before the first use: 13 x 16 bytes per thread in flight
    constexpr int NL = (NPAIR + HT_THREADS - 1) / HT_THREADS;
    float4 vv[NL];
    // steady state (block-uniform): the whole tile exists, lies past `from` and inside the producer's segment grid
    const bool steady = vrow && q_lo == 0 && q_full >= NPAIR && from_s <= 0 && seg_rel >= 0 &&
                        ((seg_rel + 2 * NPAIR - 1) >> c.seg_shift) < c.nseg;
    if (steady) {
      const float* eb = c.e + c.halo_out;
      const float na = -c.alpha;
#pragma unroll
      for (int i = 0; i < NL; i++) {
        const int q = t + i * HT_THREADS;
        if (i < NL - 1 || q < NPAIR) vv[i] = *(const float4*)(ring + ((nb32 + 2u * (unsigned)q) & rmask));
      }
#pragma unroll
      for (int i = 0; i < NL; i++) {
        const int q = t + i * HT_THREADS;
        if (i == NL - 1 && q >= NPAIR) break;
        const unsigned rel = (unsigned)(seg_rel + 2 * q);
        const float2 v0 = vrow[rel >> c.seg_shift];
        const float2 e = *(const float2*)(eb + (rel & smask));   // rel even, table 8-byte aligned
        const float ax = na * v0.x, ay = na * v0.y;
        float4 v = vv[i];
        const int pq = pad(q);
        ev[pq] = make_float2(fmaf(ax, e.x, v.x), fmaf(ay, e.x, v.y));
        od[pq] = make_float2(fmaf(ax, e.y, v.z), fmaf(ay, e.y, v.w));
      }
    } else {
#pragma unroll 1
      for (int q = t; q < NPAIR; q += HT_THREADS) {
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (q >= q_lo && q < q_full) v = *(const float4*)(ring + ((nb32 + 2u * (unsigned)q) & rmask));
        if (q == q_full && (avail & 1) && q >= q_lo) {   // odd count: only the even sample exists yet
          const float2 a = ring[(nb32 + 2u * (unsigned)q) & rmask];
          v.x = a.x; v.y = a.y;
        }
        // zero-input response of the producer's DC blocker: x -= alpha V0[segment] E[k].  The producer's segments
        // hold an even number of ring samples, so both samples of a pair share the segment.
        const int rel = seg_rel + 2 * q;
        if (vrow && 2 * q + 1 >= from_s && 2 * q < avail && rel >= 0) {
          const int sg = rel >> c.seg_shift;
          if (sg < c.nseg) {
            const float2 v0 = vrow[sg];
            const float* ep = c.e + ((unsigned)rel & smask) + c.halo_out;
            const float ax = -c.alpha * v0.x, ay = -c.alpha * v0.y;
            if (2 * q >= from_s) { const float e0 = __ldg(ep); v.x = fmaf(ax, e0, v.x); v.y = fmaf(ay, e0, v.y); }
            if (2 * q + 1 < avail) { const float e1 = __ldg(ep + 1); v.z = fmaf(ax, e1, v.z); v.w = fmaf(ay, e1, v.w); }
          }
        }
        ev[pad(q)] = make_float2(v.x, v.y);
        od[pad(q)] = make_float2(v.z, v.w);
      }
    }
  }
  __syncthreads();

  // ---- half-band decimator (A.4): out[o] = odd[o - M] + sum_j h[j] even[o - j], times scale ---------------------
  {
    const int base = (RH + PADK) * t;             // pad(RH t)
    float2 e[RH + W], hout[RH];
#pragma unroll
    for (int c = 0; c < RH + W; c++) e[c] = ev[base + HT_HH + c + PADK * ((HT_HH + c) / RH)];
#pragma unroll
    for (int r = 0; r < RH; r++) {
      const float2 o = od[base + (HT_HH + W - M + r) + PADK * ((HT_HH + W - M + r) / RH)];
      float ar = o.x, ai = o.y;
#pragma unroll
      for (int j = 0; j < 2 * M; j++) {
        ar = fmaf(p.hb[j], e[r + W - j].x, ar);
        ai = fmaf(p.hb[j], e[r + W - j].y, ai);
      }
      hout[r] = make_float2(ar * p.scale, ai * p.scale);
    }
    float2 hpre = make_float2(0.0f, 0.0f);
    if (t < HT_HH) {                           // the few outputs before the tile that the resampler window reaches
      const float2 o = od[pad(t + W - M)];
      float ar = o.x, ai = o.y;
#pragma unroll
      for (int j = 0; j < 2 * M; j++) {
        const float2 x = ev[pad(t + W - j)];
        ar = fmaf(p.hb[j], x.x, ar);
        ai = fmaf(p.hb[j], x.y, ai);
      }
      hpre = make_float2(ar * p.scale, ai * p.scale);
    }
    __syncthreads();                           // everyone has read its even samples: the buffer is reused
#pragma unroll
    for (int r = 0; r < RH; r++) hbo[base + (HT_HH + r) + PADK * ((HT_HH + r) / RH)] = hout[r];
    if (t < HT_HH) hbo[pad(t)] = hpre;
  }
  __syncthreads();

  // ---- arbitrary resampler (A.5): output j sits at input floor(j step / 2^24) with bank row (j mod P) -----------
  {
    const int base = (RH + PADK) * t;
    constexpr int NW = RH + 12;                // inputs 12 t + HH - 13 .. 12 t + HH + RH - 2
    float2 w[NW];
#pragma unroll
    for (int c = 0; c < NW; c++) w[c] = hbo[base + (HT_HH - 13 + c) + PADK * ((HT_HH - 13 + c) / RH)];
    float out[2 * HT_R];
#pragma unroll
    for (int r = 0; r < HT_R; r++) {
      const int io = Q * (r / P) + ((r % P) * Q) / P;   // input offset of output r within the thread's span
      float yr = 0.0f, yi = 0.0f;
#pragma unroll
      for (int k = 0; k < 14; k++) {
        yr = fmaf(p.arb[r % P][k], w[io + 13 - k].x, yr);
        yi = fmaf(p.arb[r % P][k], w[io + 13 - k].y, yi);
      }
      out[2 * r] = yr;
      out[2 * r + 1] = yi;
    }
    const long long j = jt + HT_R * t;
    float2* dst = p.dst + (long long)s * p.dst_stride;
    if (j >= p.j0 && j + HT_R <= p.j1) {
      // j is even: 16-byte pieces are aligned and never straddle the ring's end
#pragma unroll
      for (int r = 0; r < HT_R; r += 2)
        *(float4*)(dst + ((j + r) & p.dst_mask)) = make_float4(out[2 * r], out[2 * r + 1], out[2 * r + 2], out[2 * r + 3]);
    } else {
#pragma unroll
      for (int r = 0; r < HT_R; r++)
        if (j + r >= p.j0 && j + r < p.j1) dst[(j + r) & p.dst_mask] = make_float2(out[2 * r], out[2 * r + 1]);
    }
  }
}

template <int M, int P, int Q>
inline size_t hbarb_tile_smem() {
  constexpr int RH = HT_R / P * Q, TH = RH * HT_THREADS, W = 2 * M - 1, NPAIR = TH + HT_HH + W, NH = TH + HT_HH;
  auto pad = [](int i) { return i + ((RH % 2 == 0) ? i / RH : 0); };
  (void)NH;
  return (size_t)(2 * (pad(NPAIR) + 1)) * sizeof(float2);   // the half-band outputs reuse the even-phase buffer
}

}  // namespace pmr
