// Receiver kernels: per-channel RSSI, the squelch / channel-selector state machine, the discriminator
// of the selected channel, and the CTCSS tone detector.  They replace, for n_streams receivers at once,
//   average_power()            /root/reference/src/sdr_pmr446.c:330-336
//   find_max_rssi_channel()    :668-700
//   the state machine          :828-874
//   freqdem on the active channel :881 (one demodulator whose r_prime survives channel changes, reset on detune :866)
//   ctcss_execute()            :605-628 and ctcss_detector_analyze() :365-407.
// The state machine is one thread per stream: the host never reads device state between chunks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pmr446_b200.h"
#include "channelizer.cuh"   // fast_atan2f

namespace pmr {

constexpr int RX_TONES = PMR446_CTCSS_TONES;

// Per-stream receiver state (device memory, one element per stream).
struct RxState {
  int state, active;          // proc_scanning / proc_tuned, active channel or -1
  float rssi, ctcss_freq;
  long long n_sel;            // samples appended to the selected stream so far
  long long sel_f0, sel_f1;   // range appended by the current call
  float2 prev_cur, prev_next; // discriminator r_prime used by this call / left by it
  float dc_v1;                // CTCSS DC blocker state
  unsigned samp;              // ctcss samp_processed
  float max_power;
  int max_index, tone;
  int events;
};

// rssi[row] = 20 log10(mean |chan[row][0..ns)|), one block per (stream, channel) row
static __global__ void __launch_bounds__(128) rssi_kernel(const float2* chan, long long ld, int ns, float* rssi) {
  __shared__ float part[4];
  const float2* x = chan + (long long)blockIdx.x * ld;
  float acc = 0.0f;
  for (int k = threadIdx.x; k < ns; k += 128) {
    const float2 v = x[k];
    acc += sqrtf(fmaf(v.x, v.x, v.y * v.y));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float sum = (part[0] + part[1]) + (part[2] + part[3]);
    rssi[blockIdx.x] = 20.0f * log10f(sum / (float)ns);   // ns == 0 gives NaN like the reference's 0/0
  }
}

struct SquelchParams {
  const float* rssi;     // [S][M]
  RxState* st;
  float* u;              // [S][3][RX_TONES]: u0, u1, power
  long long* sel_range;  // [S][2] for the audio kernel
  int* sel_chan;         // [S] active channel after this chunk's update (-1 = none), for pmr446_batch_gather_channel
  int S, M, ns;
  float squelch;
  unsigned long long mask;
  int lock_max;
};

static __global__ void squelch_kernel(SquelchParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.S) return;
  RxState st = p.st[s];
  const float* r = p.rssi + (long long)s * p.M;
  // strongest enabled channel and its margin over the mean of the enabled ones
  int best = -1, enabled = 0;
  float top = 0.0f, sum = 0.0f;
  for (int i = 0; i < p.M; i++) {
    if (!((p.mask >> i) & 1ull)) continue;
    enabled++;
    const float v = r[i];
    sum += v;
    if (best < 0 || v > top) { top = v; best = i; }
  }
  const float spread = top - sum / (float)enabled;
  int ev = 0;
  bool detune = false;
  st.rssi = spread;
  if (st.state == 0) {
    if (spread > p.squelch) { st.active = best; st.state = 1; ev |= PMR446_EV_TUNED; }
  } else {
    if (p.lock_max && st.active != best) { st.active = best; ev |= PMR446_EV_CHANGED; }
    if ((double)spread < (double)p.squelch - 5.0) {
      st.active = -1;
      st.state = 0;
      st.ctcss_freq = 0.0f;
      detune = true;
      ev |= PMR446_EV_DETUNED;
    }
  }
  st.prev_cur = st.prev_next;
  if (detune) {   // freqdem_reset + ctcss_detector_reset
    st.prev_cur = st.prev_next = make_float2(0.0f, 0.0f);
    st.samp = 0;
    st.max_power = 0.0f;
    st.max_index = 0;
    st.tone = 0;
    float* u = p.u + (long long)s * 3 * RX_TONES;
    for (int j = 0; j < 3 * RX_TONES; j++) u[j] = 0.0f;
  }
  st.sel_f0 = st.n_sel;
  if (st.active >= 0) st.n_sel += p.ns;
  st.sel_f1 = st.n_sel;
  st.events = ev;
  p.sel_range[2 * s] = st.sel_f0;
  p.sel_range[2 * s + 1] = st.sel_f1;
  p.sel_chan[s] = st.active;
  p.st[s] = st;
}

// discriminator of the active channel appended to the selected-stream ring
static __global__ void rx_demod_kernel(const float2* chan, long long ld, int M, int ns, RxState* st, float ref, float* sel, long long sel_stride,
                                       long long sel_mask) {
  const int s = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int active = st[s].active;
  if (active < 0 || k >= ns) return;
  const float2* x = chan + ((long long)s * M + active) * ld;
  const float2 y = x[k];
  const float2 pv = k > 0 ? x[k - 1] : st[s].prev_cur;
  const float re = __fadd_rn(__fmul_rn(pv.x, y.x), __fmul_rn(pv.y, y.y));
  const float im = __fsub_rn(__fmul_rn(pv.x, y.y), __fmul_rn(pv.y, y.x));
  sel[(long long)s * sel_stride + ((st[s].sel_f0 + k) & sel_mask)] = atan2f(im, re) * ref;
  if (k == ns - 1) st[s].prev_next = y;
}

// Same append, from the library's discriminator ring: `row` holds the active channel's discriminator output of this call
// (pmr446_batch_gather_channel).  Only the first sample is recomputed, against the selected demodulator's own r_prime --
// it differs from the ring's value when the active channel just changed or the demodulator was reset.
static __global__ void rx_append_kernel(const float* row, long long ld, const float2* edge, int M, int ns, RxState* st, float ref, float* sel,
                                        long long sel_stride, long long sel_mask) {
  const int s = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int active = st[s].active;
  if (active < 0 || k >= ns) return;
  float v = row[(long long)s * ld + k];
  const float2* e = edge + ((long long)s * M + active) * 2;
  if (k == 0) {
    const float2 y = e[0], pv = st[s].prev_cur;
    const float re = __fadd_rn(__fmul_rn(pv.x, y.x), __fmul_rn(pv.y, y.y));
    const float im = __fsub_rn(__fmul_rn(pv.x, y.y), __fmul_rn(pv.y, y.x));
    v = fast_atan2f(im, re) * ref;
  }
  sel[(long long)s * sel_stride + ((st[s].sel_f0 + k) & sel_mask)] = v;
  if (k == ns - 1) st[s].prev_next = e[1];
}

struct CtcssParams {
  const float* lpcomp;   // [S][ld] complementary branch of this call, column = k
  long long ld;
  RxState* st;
  float* u;              // [S][3][RX_TONES]
  const float* coef;     // [RX_TONES]
  const float* freqs;    // [RX_TONES]
  float dc_a1;           // -1 + alpha
  unsigned block;
  float* ctcss_in;       // optional [S][out_ld]
  float* power_out;      // optional [S][RX_TONES]
  pmr446_rx_status* status;  // optional [S]
  long long out_ld;
};

constexpr int CT_TILE = 1024;

// One block per stream, three warps: threads j < 38 (warps 0 and 1) run tone j over the samples in order; warp 2 runs the DC
// blocker -- a serial chain of its own -- one tile AHEAD of them, in the other half of a double buffer, so a tile costs
// max(DC chain, Goertzel chains) instead of their sum (the two-warp kernel spent 43 % of its time with every tone thread
// waiting at the barrier behind the DC thread).  Same operations in the same order as before (bit-exact detector).
constexpr int CT_THREADS = 96;
__device__ __forceinline__ void ct_dc_tile(float* xs, int len, float a1, float& v1) {
  // iirfilt_rrrf DC blocker (Direct Form II): v0 = x - a1 v1 ; y = v0 - v1.  Eight samples per trip, loaded before the chain
  // starts: the shared-memory latency stays off the serial mul -> sub path
  int i = 0;
  for (; i + 8 <= len; i += 8) {
    float xv[8];
#pragma unroll
    for (int q = 0; q < 8; q++) xv[q] = xs[i + q];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const float v0 = __fsub_rn(xv[q], __fmul_rn(a1, v1));
      xs[i + q] = __fsub_rn(v0, v1);
      v1 = v0;
    }
  }
  for (; i < len; i++) {
    const float v0 = __fsub_rn(xs[i], __fmul_rn(a1, v1));
    xs[i] = __fsub_rn(v0, v1);
    v1 = v0;
  }
}
__device__ __forceinline__ void ct_tone_barrier() { asm volatile("bar.sync 1, 64;" ::: "memory"); }   // warps 0 and 1 only

static __global__ void __launch_bounds__(CT_THREADS) ctcss_kernel(CtcssParams p) {
  __shared__ float xbuf[2][CT_TILE];
  __shared__ float pw[RX_TONES];
  __shared__ float sh_v1;
  const int s = blockIdx.x, j = threadIdx.x;
  const bool dc_warp = j >= 64;
  RxState st = p.st[s];
  const int n = (int)(st.sel_f1 - st.sel_f0);
  float* ub = p.u + (long long)s * 3 * RX_TONES;
  const bool tone = j < RX_TONES;
  float u0 = tone ? ub[j] : 0.0f, u1 = tone ? ub[RX_TONES + j] : 0.0f;
  const float coef = tone ? p.coef[j] : 0.0f;
  const int had_tone = st.tone, had_code = st.max_index;
  float v1 = st.dc_v1;
  unsigned samp = st.samp;
  const float* x = p.lpcomp + (long long)s * p.ld;
  // prologue: tile 0 raw -> xbuf[0], DC on it
  {
    const int len0 = min(CT_TILE, n);
    for (int i = j; i < len0; i += CT_THREADS) xbuf[0][i] = x[i];
    __syncthreads();
    if (j == 64) ct_dc_tile(xbuf[0], len0, p.dc_a1, v1);
    __syncthreads();
  }
  for (int base = 0, k = 0; base < n; base += CT_TILE, k++) {
    float* const xs = xbuf[k & 1];
    float* const xn = xbuf[(k & 1) ^ 1];
    const int len = min(CT_TILE, n - base);
    const int len_n = min(CT_TILE, n - base - CT_TILE);           // <= 0: no next tile
    for (int i = j; i < len_n; i += CT_THREADS) xn[i] = x[base + CT_TILE + i];
    __syncthreads();
    if (dc_warp) {
      if (j == 64 && len_n > 0) ct_dc_tile(xn, len_n, p.dc_a1, v1);
    } else {
    if (p.ctcss_in)
      for (int i = j; i < len; i += 64) p.ctcss_in[(long long)s * p.out_ld + base + i] = xs[i];
    int i = 0;
    while (i < len) {
      // run to the end of the tile or of the detector block, whichever comes first (uniform over the block)
      const int run = (int)min((unsigned)(len - i), p.block - samp);
      const int e = i + run;
      for (; i + 8 <= e; i += 8) {   // same recurrence, same order; the eight broadcast loads run ahead of the serial chain
        float xv[8];
#pragma unroll
        for (int q = 0; q < 8; q++) xv[q] = xs[i + q];
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float t = u0;
          u0 = __fsub_rn(__fadd_rn(xv[q], __fmul_rn(coef, u0)), u1);
          u1 = t;
        }
      }
      for (; i < e; i++) {
        const float t = u0;
        u0 = __fsub_rn(__fadd_rn(xs[i], __fmul_rn(coef, u0)), u1);
        u1 = t;
      }
      samp += (unsigned)run;
      if (samp == p.block) {
        if (tone) {
          pw[j] = __fsub_rn(__fadd_rn(__fmul_rn(u0, u0), __fmul_rn(u1, u1)), __fmul_rn(__fmul_rn(coef, u0), u1));
          ub[2 * RX_TONES + j] = pw[j];
        }
        u0 = u1 = 0.0f;
        ct_tone_barrier();
        if (j == 0) {
          float avg = 0.0f, mx = 0.0f;
          int mi = st.max_index;
          for (int q = 0; q < RX_TONES; q++) {
            avg += pw[q];
            if (pw[q] > mx) { mx = pw[q]; mi = q; }
          }
          avg = __fdiv_rn(avg, (float)RX_TONES);
          st.max_power = mx;
          st.max_index = mi;
          st.tone = (avg > 120.0f) && (__fdiv_rn(mx, avg) > 10.0f);
        }
        ct_tone_barrier();
        samp = 0;
      }
    }
    }
    __syncthreads();   // tile k consumed, tile k + 1 filtered
  }
  if (j == 64) sh_v1 = v1;
  __syncthreads();
  v1 = sh_v1;
  if (tone) { ub[j] = u0; ub[RX_TONES + j] = u1; }
  if (p.power_out && tone) p.power_out[(long long)s * RX_TONES + j] = ub[2 * RX_TONES + j];
  if (j == 0) {
    st.dc_v1 = v1;
    st.samp = samp;
    if (st.active >= 0) {   // ctcss_execute ran for this chunk
      st.ctcss_freq = p.freqs[st.max_index];
      if (st.tone) {
        if (!had_tone) st.events |= PMR446_EV_CTCSS_ACQUIRED;
        else if (had_code != st.max_index) st.events |= PMR446_EV_CTCSS_CHANGED;
      } else if (had_tone) {
        st.events |= PMR446_EV_CTCSS_LOST;
      }
    }
    p.st[s] = st;
    if (p.status) {
      pmr446_rx_status o;
      o.state = st.state;
      o.active_chan = st.active;
      o.rssi = st.rssi;
      o.n_audio = (unsigned)n;
      o.tone_detected = st.tone;
      o.ctcss_index = st.max_index;
      o.ctcss_freq = st.ctcss_freq;
      o.max_power = st.max_power;
      o.events = st.events;
      p.status[s] = o;
    }
  }
}

}  // namespace pmr
