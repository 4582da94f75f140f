// Host-side plumbing shared by the C-ABI entry points: error reporting, device buffers,
// ring gather / history kernels, and the Frontend object that owns the DC-blocker + multi-stage
// resampler launches (the device counterpart of iirfilt_crcf + msresamp_crcf,
// /root/reference/src/sdr_pmr446.c:422-428,795-796; src/dsd_in.c:97-101,167-168).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pmr446_b200.h"
#include "design.hpp"
#include "frontend.cuh"

namespace pmr {

inline std::string& last_error_string() {
  static thread_local std::string s;
  return s;
}
inline int fail(int code, const char* msg) {
  last_error_string() = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      char buf__[256];                                                                         \
      snprintf(buf__, sizeof buf__, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return pmr::fail(PMR446_ECUDA, buf__);                                                   \
    }                                                                                          \
  } while (0)

inline int select_device(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail(PMR446_ENODEV, "no CUDA device: this library has no CPU fallback");
  if (device >= 0) {
    if (device >= count) return fail(PMR446_ENODEV, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device));
  }
  return 0;
}

inline long long next_pow2(long long v) {
  long long p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  int alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    if (cudaMalloc(&p, n) != cudaSuccess) {
      p = nullptr;
      cudaGetLastError();
      return fail(PMR446_ENOMEM, "cudaMalloc failed");
    }
    bytes = n;
    return 0;
  }
  int alloc_zero(size_t n) {
    if (int rc = alloc(n)) return rc;
    CUDA_TRY(cudaMemset(p, 0, bytes));
    return 0;
  }
  int ensure(size_t n) { return (p && bytes >= n) ? 0 : alloc(n); }
};

// out[row][k] = ring[row][(start + k) & mask], k < count
template <typename T>
__global__ void gather_ring_kernel(const T* ring, long long ring_stride, long long mask, long long start, long long count, T* out, long long out_ld) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  long long row = blockIdx.y;
  long long n = start + k;
  out[row * out_ld + k] = n >= 0 ? ring[row * ring_stride + (n & mask)] : T();
}

// dst[s][k] = x[new_base + k] taken from the two-source view (history or current chunk)
template <typename T>
__global__ void hist_update_kernel(SrcView v, T* dst, long long dst_stride_elts, long long new_base, int count) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  int s = blockIdx.y;
  long long n = new_base + k;
  T val = T();
  if (n >= v.n0) {
    if (n < v.n1) val = ((const T*)((const char*)v.cur + (long long)s * v.cur_stride))[n - v.n0];
  } else if (n >= v.hist_base) {
    val = ((const T*)((const char*)v.hist + (long long)s * v.hist_stride))[n - v.hist_base];
  }
  dst[(long long)s * dst_stride_elts + k] = val;
}

// Optional per-kernel timing with CUDA events recorded on the launch stream (bench.py's roofline).
// mark(tag) closes the interval since the previous mark and attributes it to `tag`.
enum { TM_START = 0, TM_DC = 1, TM_CASCADE0 = 2, TM_CASCADE1 = 3, TM_CASCADE2 = 4, TM_HIST = 5, TM_CHANNELIZE = 6, TM_AUDIO = 7,
       TM_WATERFALL = 8, TM_GATHER = 9, TM_NTAGS = 10 };
struct Timer {
  bool enabled = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> tag;
  size_t used = 0;
  double total_ms[TM_NTAGS] = {0};
  long long count[TM_NTAGS] = {0};
  ~Timer() { for (auto e : ev) cudaEventDestroy(e); }
  void mark(cudaStream_t st, int t) {
    if (!enabled) return;
    if (used == ev.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev.push_back(e);
      tag.push_back(0);
    }
    tag[used] = t;
    cudaEventRecord(ev[used++], st);
  }
  // folds the recorded events into the totals; call after the stream has been synchronised
  void collect() {
    for (size_t i = 1; i < used; i++) {
      if (tag[i] == TM_START) continue;
      float ms = 0.0f;
      if (cudaEventElapsedTime(&ms, ev[i - 1], ev[i]) == cudaSuccess) { total_ms[tag[i]] += ms; count[tag[i]]++; }
    }
    used = 0;
  }
  void clear() {
    used = 0;
    for (int i = 0; i < TM_NTAGS; i++) { total_ms[i] = 0; count[i] = 0; }
  }
};

typedef void (*cascade_fn)(CascadeParams);
typedef void (*dclocal_fn)(DcLocalParams);

// One cascade launch: a run of half-band stages (execution order, highest rate first), optionally
// preceded by the DC blocker and followed by the arbitrary resampler.
struct Level {
  int src = SRC_RING;
  bool dc = false, arb = false;
  int ms[4] = {0, 0, 0, 0};
  int nst = 0, D = 1, G = 16, halo = 0, seg_len = 1024;
  float hb[4][20];
  float scale = 1.0f;
  cascade_fn fn = nullptr;
  DevBuf ring;            // output ring [S][cap] float2
  long long cap = 0;
  long long n_in = 0, n_out = 0, n_hb = 0;  // absolute counts: inputs seen, outputs written, half-band outputs
  long long max_in = 0, max_out = 0;        // per chunk
};

template <int SRC, bool DC, int G, int A, int B, int C, int D, bool ARB>
static cascade_fn cfn() { return cascade_kernel<SRC, DC, G, A, B, C, D, ARB>; }

// Instantiated stage combinations; anything else is reported as unsupported.
inline cascade_fn pick_cascade(int src, bool dc, bool arb, const int* ms, int* G) {
  auto is = [&](int a, int b, int c, int d) { return ms[0] == a && ms[1] == b && ms[2] == c && ms[3] == d; };
  *G = 16;
  if (!arb) {
    if (dc && src == SRC_CU8) {
      if (is(5, 0, 0, 0)) return cfn<SRC_CU8, true, 16, 5, 0, 0, 0, false>();
      if (is(3, 5, 0, 0)) return cfn<SRC_CU8, true, 16, 3, 5, 0, 0, false>();
      if (is(3, 3, 3, 3)) return cfn<SRC_CU8, true, 16, 3, 3, 3, 3, false>();
    }
    if (dc && src == SRC_CF32) {
      if (is(5, 0, 0, 0)) return cfn<SRC_CF32, true, 16, 5, 0, 0, 0, false>();
      if (is(3, 5, 0, 0)) return cfn<SRC_CF32, true, 16, 3, 5, 0, 0, false>();
      if (is(3, 3, 3, 3)) return cfn<SRC_CF32, true, 16, 3, 3, 3, 3, false>();
    }
    if (!dc && src == SRC_CF32) {   // stand-alone msresamp_crcf (liquid shim): input is already DC-blocked
      if (is(5, 0, 0, 0)) return cfn<SRC_CF32, false, 16, 5, 0, 0, 0, false>();
      if (is(3, 5, 0, 0)) return cfn<SRC_CF32, false, 16, 3, 5, 0, 0, false>();
      if (is(3, 3, 3, 3)) return cfn<SRC_CF32, false, 16, 3, 3, 3, 3, false>();
    }
    if (!dc && src == SRC_RING) {
      if (is(5, 0, 0, 0)) return cfn<SRC_RING, false, 16, 5, 0, 0, 0, false>();
      if (is(3, 5, 0, 0)) return cfn<SRC_RING, false, 16, 3, 5, 0, 0, false>();
    }
  } else {
    if (!dc && src == SRC_RING && is(10, 0, 0, 0)) { *G = 8; return cfn<SRC_RING, false, 8, 10, 0, 0, 0, true>(); }
    if (dc && src == SRC_CU8 && is(0, 0, 0, 0)) return cfn<SRC_CU8, true, 16, 0, 0, 0, 0, true>();
    if (dc && src == SRC_CF32 && is(0, 0, 0, 0)) return cfn<SRC_CF32, true, 16, 0, 0, 0, 0, true>();
    if (!dc && src == SRC_CF32 && is(0, 0, 0, 0)) return cfn<SRC_CF32, false, 16, 0, 0, 0, 0, true>();
  }
  return nullptr;
}

struct Frontend {
  int S = 0, fmt = 0;
  bool dc = true;
  float alpha = 0.0f;
  design::MsresampPlan plan;
  std::vector<Level> levels;
  // level-0 raw history (two buffers, swapped every call)
  DevBuf hist[2];
  int hist_cur = 0;
  long long hist_base = 0;
  int hist_cap = 0;       // samples
  int bps = 8;            // bytes per input sample
  // DC blocker carry (A.1): V at (segment 0 start - halo) of the next chunk
  DevBuf v_lag, sums, v_seg;
  int max_seg0 = 0;
  DevBuf pfb;
  long long n_in = 0;     // raw samples consumed so far
  // final output (alias of levels.back())
  DevBuf out_dummy;
  DevBuf& out_ref() { return levels.back().ring; }
  struct OutView { void* p; } out;
  long long out_cap = 0, n_out = 0;
  unsigned max_chunk = 0;

  long long max_out_per_chunk() const { return levels.empty() ? 0 : levels.back().max_out; }

  int init(int n_streams, int in_fmt, float rate, float as, bool with_dc, float dc_alpha, unsigned max_chunk_, long long extra_hist) {
    S = n_streams;
    fmt = in_fmt;
    dc = with_dc;
    alpha = dc_alpha;
    max_chunk = max_chunk_;
    bps = (in_fmt == PMR446_FMT_CU8) ? 2 : 8;
    if (rate > 1.0f) return fail(PMR446_EINVAL, "front end only decimates (rate <= 1)");
    plan = design::msresamp_plan(rate, as);
    if (plan.sub_len != 14) return fail(PMR446_EINVAL, "arbitrary resampler kernel is specialised for 14 taps");
    // execution-order stage list: plan.m[stages-1] runs first
    std::vector<int> order;
    for (int g = (int)plan.stages - 1; g >= 0; g--) order.push_back((int)plan.m[g]);
    std::vector<std::vector<int>> groups;   // pre-launch groups, then the arb launch
    std::vector<int> last;
    if (!order.empty()) { last.push_back(order.back()); order.pop_back(); }
    for (size_t i = 0; i < order.size(); i += 4) groups.emplace_back(order.begin() + i, order.begin() + std::min(order.size(), i + 4));
    groups.push_back(last);   // may be empty (rate >= 0.5)
    levels.resize(groups.size());
    long long max_in = max_chunk;
    int stage_cursor = (int)plan.stages - 1;  // index into plan.m / plan.hb of the next stage to place
    for (size_t l = 0; l < groups.size(); l++) {
      Level& L = levels[l];
      L.src = (l == 0) ? (in_fmt == PMR446_FMT_CU8 ? SRC_CU8 : SRC_CF32) : SRC_RING;
      L.dc = (l == 0) && dc;
      L.arb = (l + 1 == groups.size());
      L.nst = (int)groups[l].size();
      L.D = 1 << L.nst;
      memset(L.hb, 0, sizeof L.hb);
      int halo = 0;
      for (int k = 0; k < L.nst; k++) {
        L.ms[k] = groups[l][k];
        const std::vector<float>& t = plan.hb[stage_cursor];
        for (size_t j = 0; j < t.size(); j++) L.hb[k][j] = t[j];
        halo += (4 * L.ms[k] - 2 + 1) << k;
        stage_cursor--;
      }
      if (L.arb) halo += 14 * L.D;
      L.scale = 1.0f / (float)L.D;
      L.fn = pick_cascade(L.src, L.dc, L.arb, L.ms, &L.G);
      if (!L.fn) {
        char msg[160];
        snprintf(msg, sizeof msg, "resampler plan not built: level %zu src=%d dc=%d arb=%d stages=[%d,%d,%d,%d]", l, L.src, (int)L.dc,
                 (int)L.arb, L.ms[0], L.ms[1], L.ms[2], L.ms[3]);
        return fail(PMR446_EINVAL, msg);
      }
      const int unit = std::max(16, L.D);   // segment granularity
      L.halo = (halo + L.D + unit - 1) / unit * unit;
      // segment length: as long as possible while keeping >= ~150k threads in flight
      int seg = 4096;
      const int seg_min = std::max(256, ((4 * L.halo + unit - 1) / unit) * unit);
      while (seg > seg_min && (long long)S * ((max_in + seg - 1) / seg) < 148LL * 1024) seg >>= 1;
      L.seg_len = std::max(seg, seg_min);
      L.max_in = max_in;
      long long max_hb = max_in / L.D + 1;
      L.max_out = L.arb ? (long long)design::arb_outputs_after((uint64_t)max_hb, plan.step) + 2 : max_hb;
      long long need = L.max_out + 64 + (l + 1 < groups.size() ? 0 : extra_hist);
      if (l + 1 < groups.size()) need += 4096;  // next level's halo (set below once known; generous bound)
      L.cap = next_pow2(need);
      if (int rc = L.ring.alloc_zero((size_t)S * L.cap * sizeof(float2))) return rc;
      max_in = L.max_out;
    }
    for (size_t l = 1; l < levels.size(); l++)
      if (levels[l].halo + 64 > 4096) return fail(PMR446_EINVAL, "internal: halo exceeds ring slack");
    // arbitrary resampler bank, rows padded to 16 floats
    std::vector<float> rows((size_t)plan.npfb * 16, 0.0f);
    for (unsigned i = 0; i < plan.npfb; i++)
      for (unsigned k = 0; k < plan.sub_len; k++) rows[(size_t)i * 16 + k] = plan.pfb[(size_t)i * plan.sub_len + k];
    if (int rc = pfb.alloc(rows.size() * sizeof(float))) return rc;
    CUDA_TRY(cudaMemcpy(pfb.p, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice));
    // raw history and DC carry
    Level& L0 = levels[0];
    hist_cap = L0.halo + 16;
    for (int i = 0; i < 2; i++)
      if (int rc = hist[i].alloc_zero((size_t)S * hist_cap * bps)) return rc;
    max_seg0 = (int)((max_chunk + 32 + L0.seg_len - 1) / L0.seg_len) + 2;
    if (int rc = v_lag.alloc_zero((size_t)S * sizeof(float2))) return rc;
    if (int rc = sums.alloc_zero((size_t)S * max_seg0 * sizeof(float2))) return rc;
    if (int rc = v_seg.alloc_zero((size_t)S * max_seg0 * sizeof(float2))) return rc;
    reset_counters();
    out.p = levels.back().ring.p;
    out_cap = levels.back().cap;
    return 0;
  }

  void reset_counters() {
    n_in = 0;
    n_out = 0;
    hist_cur = 0;
    hist_base = -(long long)levels[0].halo;
    for (auto& L : levels) L.n_in = L.n_out = L.n_hb = 0;
  }
  void reset() {
    reset_counters();
    for (auto& L : levels) cudaMemset(L.ring.p, 0, L.ring.bytes);
    for (int i = 0; i < 2; i++) cudaMemset(hist[i].p, 0, hist[i].bytes);
    cudaMemset(v_lag.p, 0, v_lag.bytes);
  }

  // Consumes n raw samples per stream ([n_in, n_in + n)); afterwards n_out is the absolute
  // number of output samples in the `out` ring.
  int execute(const void* iq, long long iq_stride, unsigned n, cudaStream_t st, int* launches, Timer* tm = nullptr) {
    const long long n0 = n_in, n1 = n_in + n;
    Timer dummy;
    if (!tm) tm = &dummy;
    SrcView v0;
    v0.hist = hist[hist_cur].p;
    v0.hist_stride = (long long)hist_cap * bps;
    v0.hist_base = hist_base;
    v0.cur = iq;
    v0.cur_stride = iq_stride;
    v0.n0 = n0;
    v0.n1 = n1;
    v0.ring_mask = 0;
    v0.cur_aligned = (n0 % 16 == 0) && (((uintptr_t)iq) % 16 == 0) && (iq_stride % 16 == 0);

    long long lvl_n0 = n0, lvl_n1 = n1;
    for (size_t l = 0; l < levels.size(); l++) {
      Level& L = levels[l];
      const int unit = std::max(L.G, L.D);
      const long long seg0 = lvl_n0 / unit * unit, seg0_next = lvl_n1 / unit * unit;
      const long long out0 = lvl_n0 / L.D, out1 = lvl_n1 / L.D;
      long long span = (out1 > out0) ? out1 * L.D - seg0 : seg0_next - seg0;
      const int nseg = (int)((span + L.seg_len - 1) / L.seg_len);
      SrcView sv;
      if (l == 0) {
        sv = v0;
      } else {
        const Level& P = levels[l - 1];
        sv.hist = nullptr;
        sv.hist_stride = 0;
        sv.hist_base = 0;
        sv.cur = P.ring.p;
        sv.cur_stride = P.cap * (long long)sizeof(float2);
        sv.n0 = lvl_n0;
        sv.n1 = lvl_n1;
        sv.ring_mask = P.cap - 1;
        sv.cur_aligned = 1;
      }
      if (L.dc && nseg > 0) {
        if (nseg > max_seg0) return fail(PMR446_ERANGE, "internal: segment count exceeds allocation");
        DcLocalParams dp;
        dp.src = sv;
        dp.n_streams = S;
        dp.nseg = nseg;
        dp.p0 = seg0 - L.halo;
        dp.seg_len = L.seg_len;
        dp.end = seg0_next - L.halo;
        dp.c = 1.0f - alpha;
        dp.sums = (float2*)sums.p;
        const long long threads = (long long)S * nseg;
        if (L.src == SRC_CU8) dc_local_kernel<SRC_CU8><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(dp);
        else dc_local_kernel<SRC_CF32><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(dp);
        DcScanParams sp;
        sp.n_streams = S;
        sp.nseg = nseg;
        sp.sums = (const float2*)sums.p;
        sp.v_seg = (float2*)v_seg.p;
        sp.v_lag = (float2*)v_lag.p;
        sp.p0 = dp.p0;
        sp.seg_len = L.seg_len;
        sp.end = dp.end;
        sp.c = dp.c;
        sp.decay_full = (float)pow((double)dp.c, (double)L.seg_len);
        dc_scan_kernel<<<(S + 127) / 128, 128, 0, st>>>(sp);
        *launches += 2;
        tm->mark(st, TM_DC);
      }
      long long new_out = L.n_out;
      if (out1 > out0) {
        CascadeParams cp;
        cp.src = sv;
        cp.n_streams = S;
        cp.nseg = nseg;
        cp.seg0 = seg0;
        cp.seg_len = L.seg_len;
        cp.halo = L.halo;
        cp.out0 = out0;
        cp.out1 = out1;
        cp.scale = L.scale;
        // liquid stores a1 = -1 + alpha rounded to float32 and computes y = v[n] - v[n-1] with
        // v[n] = x - a1 v[n-1]; the effective feedback is therefore 1 - fl(1 - alpha), not alpha.
        const float c_pole = 1.0f - alpha;
        cp.alpha = 1.0f - c_pole;
        cp.v_seg = (const float2*)v_seg.p;
        cp.step = plan.step;
        cp.bits = (int)plan.bits;
        cp.pfb = (const float*)pfb.p;
        cp.dst = (float2*)L.ring.p;
        cp.dst_stride = L.cap;
        cp.dst_mask = L.cap - 1;
        memcpy(cp.hb, L.hb, sizeof cp.hb);
        const long long threads = (long long)S * nseg;
        L.fn<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(cp);
        *launches += 1;
        tm->mark(st, TM_CASCADE0 + (int)std::min<size_t>(l, 2));
        new_out = L.arb ? (long long)design::arb_outputs_after((uint64_t)out1, plan.step) : out1;
      }
      L.n_in = lvl_n1;
      L.n_hb = out1;
      lvl_n0 = L.n_out;
      L.n_out = new_out;
      lvl_n1 = new_out;
    }
    // save the raw tail for the next call
    {
      const Level& L0 = levels[0];
      const int unit = std::max(L0.G, L0.D);
      const long long new_base = n1 / unit * unit - L0.halo;
      const int count = (int)(n1 - new_base);
      if (count > hist_cap) return fail(PMR446_ERANGE, "internal: history overflow");
      DevBuf& dst = hist[hist_cur ^ 1];
      dim3 grid((count + 127) / 128, S);
      if (count > 0) {
        if (bps == 2) hist_update_kernel<uint16_t><<<grid, 128, 0, st>>>(v0, (uint16_t*)dst.p, hist_cap, new_base, count);
        else hist_update_kernel<float2><<<grid, 128, 0, st>>>(v0, (float2*)dst.p, hist_cap, new_base, count);
        *launches += 1;
        tm->mark(st, TM_HIST);
      }
      hist_cur ^= 1;
      hist_base = new_base;
    }
    n_in = n1;
    n_out = levels.back().n_out;
    return 0;
  }
};

}  // namespace pmr
