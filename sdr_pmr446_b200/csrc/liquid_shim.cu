// liquid-dsp-signature shim (include/pmr446_liquid_shim.h): single-stream objects whose block
// calls run on the GPU.  Call sites served: /root/reference/src/sdr_pmr446.c:420-518,795-913 and
// /root/reference/src/dsd_in.c:95-124,167-170.  Algorithms: SURVEY.md Appendix A.
//
// The multi-stage resamplers and the waterfall reuse the batched kernels (Frontend, dsd.cuh,
// spectrum.cuh) with n_streams = 1; the small objects (first-order IIR, FIR, discriminator, NCO,
// analysis filter bank) have simple dedicated kernels below.  Throughput lives in the coarse
// batched API (pmr446_b200.h); this tier exists for literal source compatibility.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/pmr446_liquid_shim.h"
#include "common_host.hpp"
#include "frontend_host.hpp"
#include "design.hpp"
#include "dsd.cuh"
#include "spectrum.cuh"

using namespace pmr;

namespace {

struct Scratch {   // grow-only device staging shared by the objects of one thread
  DevBuf in, out;
};
Scratch& scratch() {
  static thread_local Scratch s;
  return s;
}
bool have_device() { return select_device(-1) == 0; }

// Small per-frame calls (nco_crcf_mix_block_down on 16 samples, firpfbch_crcf_analyzer_execute) are dominated by the
// two synchronous copies around the kernel.  A page-locked, device-mapped staging pair lets the kernel read its input
// and write its result over PCIe directly, and a completion word in the same kind of memory replaces the driver
// synchronisation: one launch + a short spin per call.
struct MappedPair {
  void *in = nullptr, *out = nullptr;
  size_t in_bytes = 0, out_bytes = 0;
  ~MappedPair() {
    if (in) cudaFreeHost(in);
    if (out) cudaFreeHost(out);
    if (flag) cudaFreeHost((void*)flag);
  }
  static int grow(void** p, size_t* have, size_t need) {
    if (*p && *have >= need) return 0;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    size_t n = need < 4096 ? 4096 : need;
    if (cudaHostAlloc(p, n, cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); *p = nullptr; *have = 0; return -1; }
    *have = n;
    return 0;
  }
  volatile unsigned* flag = nullptr;   // completion word written by single-block kernels
  unsigned seq = 0;
  int ensure(size_t in_need, size_t out_need) {
    if (!flag) {
      void* f = nullptr;
      if (cudaHostAlloc(&f, 64, cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return -1; }
      flag = (volatile unsigned*)f;
      *flag = 0;
    }
    return grow(&in, &in_bytes, in_need) || grow(&out, &out_bytes, out_need);
  }
  // waits for the kernel launched with (flag, seq): spins on the mapped word for a bounded time, then falls back to
  // a stream synchronisation (which also surfaces launch / execution errors)
  int wait() {
    if (cudaPeekAtLastError() == cudaSuccess) {
      for (int spin = 0; spin < 200000; spin++)
        if (*flag == seq) return 0;
    }
    return cudaStreamSynchronize(0) == cudaSuccess ? 0 : -1;
  }
};
MappedPair& mapped() {
  static thread_local MappedPair m;
  return m;
}
constexpr size_t kMappedMax = 64 * 1024;   // larger blocks amortise the copies and use device staging

// ---------------------------------------------------------------------------- first-order IIR
// v[n] = x[n] + c v[n-1];  y[n] = b0 v[n] + b1 v[n-1]   (Direct Form II, A.1)
constexpr int IIR_SEG = 256;
template <typename T> __device__ __forceinline__ T t_zero();
template <> __device__ __forceinline__ float t_zero<float>() { return 0.0f; }
template <> __device__ __forceinline__ float2 t_zero<float2>() { return make_float2(0.0f, 0.0f); }
__device__ __forceinline__ float axpy(float a, float x, float y) { return fmaf(a, x, y); }
__device__ __forceinline__ float2 axpy(float a, float2 x, float2 y) { return make_float2(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y)); }
__device__ __forceinline__ float comb(float b0, float v, float b1, float v1) { return fmaf(b0, v, b1 * v1); }
__device__ __forceinline__ float2 comb(float b0, float2 v, float b1, float2 v1) {
  return make_float2(fmaf(b0, v.x, b1 * v1.x), fmaf(b0, v.y, b1 * v1.y));
}

template <typename T>
__global__ void iir1_local_kernel(const T* x, int n, float c, T* sums) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int a = t * IIR_SEG;
  if (a >= n) return;
  int b = min(a + IIR_SEG, n);
  T s = t_zero<T>();
  for (int i = a; i < b; i++) s = axpy(c, s, x[i]);
  sums[t] = s;
}
template <typename T>
__global__ void iir1_apply_kernel(const T* x, T* y, int n, float c, float b0, float b1, float decay_seg, const T* sums, T* state) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int a = t * IIR_SEG;
  if (a >= n) return;
  int b = min(a + IIR_SEG, n);
  T v = state[0];
  for (int u = 0; u < t; u++) v = axpy(decay_seg, v, sums[u]);   // V at the segment start
  for (int i = a; i < b; i++) {
    T vn = axpy(c, v, x[i]);
    y[i] = comb(b0, vn, b1, v);
    v = vn;
  }
  if (b == n) state[1] = v;   // picked up by the host after the launch (state[0] stays valid for the other threads)
}

template <typename T>
struct Iir1 {
  float c = 0, b0 = 1, b1 = 0;
  DevBuf state, sums;
  int init(float c_, float b0_, float b1_) {
    c = c_; b0 = b0_; b1 = b1_;
    return state.alloc_zero(2 * sizeof(T));
  }
  int run(const T* hx, unsigned n, T* hy) {
    if (n == 0) return 0;
    Scratch& sc = scratch();
    const int nseg = (int)((n + IIR_SEG - 1) / IIR_SEG);
    if (sc.in.ensure(n * sizeof(T)) || sc.out.ensure(n * sizeof(T)) || sums.ensure((size_t)nseg * sizeof(T))) return LIQUID_EICONFIG;
    CUDA_TRY(cudaMemcpy(sc.in.p, hx, n * sizeof(T), cudaMemcpyHostToDevice));
    iir1_local_kernel<T><<<(nseg + 127) / 128, 128>>>((const T*)sc.in.p, (int)n, c, (T*)sums.p);
    iir1_apply_kernel<T><<<(nseg + 127) / 128, 128>>>((const T*)sc.in.p, (T*)sc.out.p, (int)n, c, b0, b1, powf(c, (float)IIR_SEG),
                                                      (const T*)sums.p, (T*)state.p);
    CUDA_TRY(cudaMemcpy(hy, sc.out.p, n * sizeof(T), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(state.p, (const T*)state.p + 1, sizeof(T), cudaMemcpyDeviceToDevice));
    return 0;
  }
};

// ---------------------------------------------------------------------------- FIR (A.10)
__global__ void fir_kernel(const float* hx /*[hist(nh) | x(n)]*/, int nh, int n, const float* h, int nt, float* y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = hx + nh + i;
  float acc = 0.0f;
  for (int k = 0; k < nt; k++) acc = fmaf(h[k], p[-k], acc);
  y[i] = acc;
}

// ---------------------------------------------------------------------------- discriminator (A.9)
__global__ void freqdem_kernel(const float2* x /*[prev | x(n)]*/, int n, float ref, float* m) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 p = x[i], c = x[i + 1];
  const float re = __fadd_rn(__fmul_rn(p.x, c.x), __fmul_rn(p.y, c.y));
  const float im = __fsub_rn(__fmul_rn(p.x, c.y), __fmul_rn(p.y, c.x));
  m[i] = atan2f(im, re) * ref;
}

// ---------------------------------------------------------------------------- NCO mix-down (A.7)
// `done` (optional, single-block launches only): a word in mapped host memory that receives `seq` once every result is
// visible to the host, so the caller can wait on it instead of paying a driver synchronisation.
__device__ __forceinline__ void signal_host(volatile unsigned* done, unsigned seq) {
  __threadfence_system();
  __syncthreads();
  if (done && threadIdx.x == 0) *done = seq;
}
__global__ void nco_mix_kernel(const float2* x, float2* y, int n, unsigned theta0, unsigned dtheta, volatile unsigned* done = nullptr,
                               unsigned seq = 0) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const unsigned th = theta0 + (unsigned)i * dtheta;
    float sn, cs;
    sincospif((float)(int)th * (1.0f / 2147483648.0f), &sn, &cs);
    const float2 v = x[i];
    y[i] = make_float2(fmaf(v.x, cs, v.y * sn), fmaf(v.y, cs, -v.x * sn));
  }
  if (done) signal_host(done, seq);
}

// ---------------------------------------------------------------------------- analysis filter bank, one frame (A.8)
// windows: [M][p] ring per branch (slot `pos` is the newest after the push).  Generic M (direct DFT).
__global__ void pfbch_frame_kernel(const float2* x, float2* win, const float* taps, int M, int p, int pos, float2* y,
                                   volatile unsigned* done = nullptr, unsigned seq = 0) {
  extern __shared__ float2 X[];
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    // commutator: x[k] goes to branch M-1-k  (filter_index starts at M-1 and decrements)
    float2* w = win + (size_t)i * p;
    w[pos] = x[M - 1 - i];
    float ar = 0.0f, ai = 0.0f;
    for (int n = 0; n < p; n++) {
      const float2 v = w[(pos - n + p) % p];
      const float h = taps[(size_t)i * p + n];
      ar = fmaf(h, v.x, ar);
      ai = fmaf(h, v.y, ai);
    }
    X[M - 1 - i] = make_float2(ar, ai);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < M; c += blockDim.x) {
    float yr = 0.0f, yi = 0.0f;
    for (int n = 0; n < M; n++) {
      float sn, cs;
      sincospif(-2.0f * (float)((c * n) % M) / (float)M, &sn, &cs);
      yr += X[n].x * cs - X[n].y * sn;
      yi += X[n].x * sn + X[n].y * cs;
    }
    y[c] = make_float2(yr, yi);
  }
  if (done) signal_host(done, seq);
}

}  // namespace

// ============================================================================ iirfilt
struct iirfilt_crcf_s { Iir1<float2> f; };
struct iirfilt_rrrf_s { Iir1<float> f; };

extern "C" iirfilt_crcf iirfilt_crcf_create_dc_blocker(float alpha) {
  if (!have_device()) return nullptr;
  iirfilt_crcf q = new iirfilt_crcf_s();
  const float a1 = -1.0f + alpha;   // liquid: b = {1, -1}, a = {1, -1 + alpha}
  if (q->f.init(-a1, 1.0f, -1.0f)) { delete q; return nullptr; }
  return q;
}
extern "C" int iirfilt_crcf_execute_block(iirfilt_crcf q, liquid_float_complex* x, unsigned int n, liquid_float_complex* y) {
  if (!q || (!x && n) || (!y && n)) return LIQUID_EICONFIG;
  return q->f.run((const float2*)x, n, (float2*)y) ? LIQUID_EICONFIG : LIQUID_OK;
}
extern "C" int iirfilt_crcf_destroy(iirfilt_crcf q) { delete q; return LIQUID_OK; }

extern "C" iirfilt_rrrf iirfilt_rrrf_create(float* b, unsigned int nb, float* a, unsigned int na) {
  if (!b || !a || nb == 0 || na == 0 || nb > 2 || na > 2 || a[0] == 0.0f) return nullptr;   // the reference only builds order <= 1
  if (!have_device()) return nullptr;
  iirfilt_rrrf q = new iirfilt_rrrf_s();
  const float a0 = a[0];
  const float a1 = na > 1 ? a[1] / a0 : 0.0f, b0 = b[0] / a0, b1 = nb > 1 ? b[1] / a0 : 0.0f;
  if (q->f.init(-a1, b0, b1)) { delete q; return nullptr; }
  return q;
}
extern "C" iirfilt_rrrf iirfilt_rrrf_create_dc_blocker(float alpha) {
  float b[2] = {1.0f, -1.0f}, a[2] = {1.0f, -1.0f + alpha};
  return iirfilt_rrrf_create(b, 2, a, 2);
}
extern "C" int iirfilt_rrrf_execute_block(iirfilt_rrrf q, float* x, unsigned int n, float* y) {
  if (!q || (!x && n) || (!y && n)) return LIQUID_EICONFIG;
  return q->f.run(x, n, y) ? LIQUID_EICONFIG : LIQUID_OK;
}
extern "C" int iirfilt_rrrf_destroy(iirfilt_rrrf q) { delete q; return LIQUID_OK; }

// ============================================================================ msresamp_crcf (decimating)
struct msresamp_crcf_s {
  Frontend fe;
  float rate, as;
  unsigned max_chunk;
  DevBuf d_in, d_out;
};
static int msresamp_crcf_setup(msresamp_crcf q, unsigned max_chunk) {
  q->max_chunk = max_chunk;
  return q->fe.init(1, PMR446_FMT_CF32, q->rate, q->as, false, 0.0f, max_chunk, 64);
}
extern "C" msresamp_crcf msresamp_crcf_create(float rate, float as) {
  if (!(rate > 0.0f) || !have_device()) return nullptr;
  msresamp_crcf q = new msresamp_crcf_s();
  q->rate = rate;
  q->as = as;
  if (msresamp_crcf_setup(q, 262144)) { delete q; return nullptr; }
  return q;
}
extern "C" int msresamp_crcf_execute(msresamp_crcf q, liquid_float_complex* x, unsigned int nx, liquid_float_complex* y, unsigned int* ny) {
  if (!q || !ny || (!x && nx)) return LIQUID_EICONFIG;
  unsigned done = 0, produced = 0;
  *ny = 0;
  while (done < nx) {   // larger blocks than the staging size are fed in pieces
    const unsigned n = std::min(nx - done, q->max_chunk);
    if (q->d_in.ensure((size_t)q->max_chunk * 8)) return LIQUID_EICONFIG;
    if (n) CUDA_TRY(cudaMemcpy(q->d_in.p, (const float2*)x + done, (size_t)n * 8, cudaMemcpyHostToDevice));
    const long long r0 = q->fe.n_out;
    int launches = 0;
    if (q->fe.execute(q->d_in.p, (long long)q->max_chunk * 8, n, nullptr, &launches)) return LIQUID_EICONFIG;
    const long long cnt = q->fe.n_out - r0;
    if (cnt > 0) {
      if (q->d_out.ensure((size_t)cnt * 8)) return LIQUID_EICONFIG;
      gather_ring_kernel<float2><<<dim3((unsigned)((cnt + 255) / 256), 1), 256>>>((const float2*)q->fe.out.p, q->fe.out_cap, q->fe.out_cap - 1, r0, cnt,
                                                                                  (float2*)q->d_out.p, cnt);
      CUDA_TRY(cudaMemcpy((float2*)y + produced, q->d_out.p, (size_t)cnt * 8, cudaMemcpyDeviceToHost));
    } else {
      CUDA_TRY(cudaDeviceSynchronize());
    }
    produced += (unsigned)cnt;
    done += n;
  }
  *ny = produced;
  return LIQUID_OK;
}
extern "C" int msresamp_crcf_print(msresamp_crcf q) {
  if (!q) return LIQUID_EICONFIG;
  const design::MsresampPlan& p = q->fe.plan;
  printf("<pmr446_b200 msresamp_crcf rate=%g, halfband stages=%u (m:", p.rate, p.stages);
  for (unsigned i = 0; i < p.stages; i++) printf(" %u", p.m[i]);
  printf("), arbitrary rate=%g step=%u npfb=%u, %zu CUDA launch level(s)>\n", p.rate_arb, p.step, p.npfb, q->fe.levels.size());
  return LIQUID_OK;
}
extern "C" int msresamp_crcf_destroy(msresamp_crcf q) {
  if (q) cudaDeviceSynchronize();
  delete q;
  return LIQUID_OK;
}

// ============================================================================ msresamp_rrrf (interpolating, src/dsd_in.c:104)
struct msresamp_rrrf_s {
  design::MsresampPlan up;
  DevBuf fm, z, pfb, d_out;
  long long fm_cap = 0, z_cap = 0, n_in = 0, n_z = 0;
  unsigned max_chunk = 65536;
};
extern "C" msresamp_rrrf msresamp_rrrf_create(float rate, float as) {
  if (!(rate > 0.0f) || !have_device()) return nullptr;
  msresamp_rrrf q = new msresamp_rrrf_s();
  q->up = design::msresamp_plan(rate, as);
  if (!q->up.interp || q->up.stages != 1 || q->up.sub_len != 14) {
    fail(PMR446_EINVAL, "msresamp_rrrf shim supports interpolation rates in (2, 4] (one half-band stage)");
    delete q;
    return nullptr;
  }
  q->fm_cap = next_pow2(q->max_chunk + 64);
  q->z_cap = next_pow2(2 * (long long)q->max_chunk + 64);
  std::vector<float> rows((size_t)q->up.npfb * 16, 0.0f);
  for (unsigned i = 0; i < q->up.npfb; i++)
    for (unsigned k = 0; k < 14; k++) rows[(size_t)i * 16 + k] = q->up.pfb[(size_t)i * 14 + k];
  if (q->fm.alloc_zero(q->fm_cap * 4) || q->z.alloc_zero(q->z_cap * 4) || q->pfb.alloc(rows.size() * 4)) { delete q; return nullptr; }
  cudaMemcpy(q->pfb.p, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice);
  return q;
}
extern "C" int msresamp_rrrf_execute(msresamp_rrrf q, float* x, unsigned int nx, float* y, unsigned int* ny) {
  if (!q || !ny || (!x && nx)) return LIQUID_EICONFIG;
  unsigned done = 0, produced = 0;
  while (done < nx) {
    const unsigned n = std::min(nx - done, q->max_chunk);
    // append to the input ring (may wrap)
    const long long pos = q->n_in & (q->fm_cap - 1), first = std::min<long long>(n, q->fm_cap - pos);
    CUDA_TRY(cudaMemcpy((float*)q->fm.p + pos, x + done, first * 4, cudaMemcpyHostToDevice));
    if (first < n) CUDA_TRY(cudaMemcpy(q->fm.p, x + done + first, (n - first) * 4, cudaMemcpyHostToDevice));
    q->n_in += n;
    const long long k0 = q->n_z, k1 = (long long)design::arb_outputs_after((uint64_t)q->n_in, q->up.step);
    if (k1 > k0) {
      const long long cnt = 2 * (k1 - k0);
      if (q->d_out.ensure((size_t)cnt * 4)) return LIQUID_EICONFIG;
      dim3 g((unsigned)((k1 - k0 + 255) / 256), 1);
      dsd_arb_kernel<<<g, 256>>>((const float*)q->fm.p, q->fm_cap, q->fm_cap - 1, (float*)q->z.p, q->z_cap, q->z_cap - 1, k0, k1, q->up.step,
                                 (int)q->up.bits, (const float*)q->pfb.p);
      DsdInterpParams ip;
      ip.z = (const float*)q->z.p;
      ip.z_stride = q->z_cap;
      ip.z_mask = q->z_cap - 1;
      ip.k0 = k0;
      ip.k1 = k1;
      ip.m = (int)q->up.m[0];
      memset(ip.hb, 0, sizeof ip.hb);
      for (size_t j = 0; j < q->up.hb[0].size(); j++) ip.hb[j] = q->up.hb[0][j];
      ip.audio = (float*)q->d_out.p;
      ip.pcm = nullptr;
      ip.out_ld = cnt;
      dsd_interp_kernel<<<g, 256>>>(ip);
      CUDA_TRY(cudaMemcpy(y + produced, q->d_out.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
      produced += (unsigned)cnt;
    }
    q->n_z = k1;
    done += n;
  }
  *ny = produced;
  return LIQUID_OK;
}
extern "C" int msresamp_rrrf_print(msresamp_rrrf q) {
  if (!q) return LIQUID_EICONFIG;
  printf("<pmr446_b200 msresamp_rrrf rate=%g, arbitrary rate=%g step=%u, halfband m=%u>\n", q->up.rate, q->up.rate_arb, q->up.step, q->up.m[0]);
  return LIQUID_OK;
}
extern "C" int msresamp_rrrf_destroy(msresamp_rrrf q) {
  if (q) cudaDeviceSynchronize();
  delete q;
  return LIQUID_OK;
}

// ============================================================================ nco
struct nco_crcf_s { unsigned theta = 0, d_theta = 0; };
extern "C" nco_crcf nco_crcf_create(liquid_ncotype) { return have_device() ? new nco_crcf_s() : nullptr; }
extern "C" int nco_crcf_set_frequency(nco_crcf q, float dtheta) {
  if (!q) return LIQUID_EICONFIG;
  q->d_theta = design::nco_dtheta(dtheta);
  return LIQUID_OK;
}
extern "C" int nco_crcf_step(nco_crcf q) {
  if (!q) return LIQUID_EICONFIG;
  q->theta += q->d_theta;   // phase bookkeeping only; the rotation itself runs on the device
  return LIQUID_OK;
}
static int nco_mix(nco_crcf q, const float2* x, float2* y, unsigned n, unsigned dtheta) {
  if ((size_t)n * 8 <= kMappedMax) {
    MappedPair& m = mapped();
    if (m.ensure((size_t)n * 8, (size_t)n * 8)) return LIQUID_EICONFIG;
    memcpy(m.in, x, (size_t)n * 8);
    if (n <= 256) {
      nco_mix_kernel<<<1, 256>>>((const float2*)m.in, (float2*)m.out, (int)n, q->theta, dtheta, m.flag, ++m.seq);
      if (m.wait()) return LIQUID_EICONFIG;
    } else {
      nco_mix_kernel<<<(n + 255) / 256, 256>>>((const float2*)m.in, (float2*)m.out, (int)n, q->theta, dtheta);
      CUDA_TRY(cudaStreamSynchronize(0));
    }
    memcpy(y, m.out, (size_t)n * 8);
    return LIQUID_OK;
  }
  Scratch& sc = scratch();
  if (sc.in.ensure((size_t)n * 8) || sc.out.ensure((size_t)n * 8)) return LIQUID_EICONFIG;
  CUDA_TRY(cudaMemcpy(sc.in.p, x, (size_t)n * 8, cudaMemcpyHostToDevice));
  nco_mix_kernel<<<(n + 255) / 256, 256>>>((const float2*)sc.in.p, (float2*)sc.out.p, (int)n, q->theta, dtheta);
  CUDA_TRY(cudaMemcpy(y, sc.out.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return LIQUID_OK;
}
extern "C" int nco_crcf_mix_down(nco_crcf q, liquid_float_complex x, liquid_float_complex* y) {
  if (!q || !y) return LIQUID_EICONFIG;
  return nco_mix(q, (const float2*)&x, (float2*)y, 1, 0);
}
extern "C" int nco_crcf_mix_block_down(nco_crcf q, liquid_float_complex* x, liquid_float_complex* y, unsigned int n) {
  if (!q || (!x && n) || (!y && n)) return LIQUID_EICONFIG;
  if (n == 0) return LIQUID_OK;
  int rc = nco_mix(q, (const float2*)x, (float2*)y, n, q->d_theta);
  q->theta += n * q->d_theta;
  return rc;
}
extern "C" int nco_crcf_destroy(nco_crcf q) { delete q; return LIQUID_OK; }

// ============================================================================ firpfbch analyzer
struct firpfbch_crcf_s {
  unsigned M = 0, p = 0;
  int pos = -1;
  DevBuf win, taps, x, y;
};
extern "C" firpfbch_crcf firpfbch_crcf_create_kaiser(int type, unsigned int M, unsigned int m, float as) {
  if (type != LIQUID_ANALYZER || M == 0 || m == 0 || M > 4096 || !have_device()) return nullptr;
  firpfbch_crcf q = new firpfbch_crcf_s();
  q->M = M;
  q->p = 2 * m;
  std::vector<float> t = design::pfbch_taps(M, m, as);
  if (q->win.alloc_zero((size_t)M * q->p * 8) || q->taps.alloc(t.size() * 4) || q->x.alloc((size_t)M * 8) || q->y.alloc((size_t)M * 8)) {
    delete q;
    return nullptr;
  }
  cudaMemcpy(q->taps.p, t.data(), t.size() * 4, cudaMemcpyHostToDevice);
  return q;
}
extern "C" int firpfbch_crcf_analyzer_execute(firpfbch_crcf q, liquid_float_complex* x, liquid_float_complex* y) {
  if (!q || !x || !y) return LIQUID_EICONFIG;
  const size_t bytes = (size_t)q->M * 8;
  q->pos = (q->pos + 1) % (int)q->p;
  if (bytes <= kMappedMax) {
    MappedPair& m = mapped();
    if (m.ensure(bytes, bytes)) return LIQUID_EICONFIG;
    memcpy(m.in, x, bytes);
    pfbch_frame_kernel<<<1, 256, bytes>>>((const float2*)m.in, (float2*)q->win.p, (const float*)q->taps.p, (int)q->M, (int)q->p, q->pos,
                                          (float2*)m.out, m.flag, ++m.seq);
    if (m.wait()) return LIQUID_EICONFIG;
    memcpy(y, m.out, bytes);
    return LIQUID_OK;
  }
  CUDA_TRY(cudaMemcpy(q->x.p, x, bytes, cudaMemcpyHostToDevice));
  pfbch_frame_kernel<<<1, 256, bytes>>>((const float2*)q->x.p, (float2*)q->win.p, (const float*)q->taps.p, (int)q->M, (int)q->p, q->pos,
                                        (float2*)q->y.p);
  CUDA_TRY(cudaMemcpy(y, q->y.p, bytes, cudaMemcpyDeviceToHost));
  return LIQUID_OK;
}
extern "C" int firpfbch_crcf_destroy(firpfbch_crcf q) { delete q; return LIQUID_OK; }

// ============================================================================ freqdem
struct freqdem_s {
  float kf = 0.5f, ref = 1.0f;
  float2 r_prime = {0.0f, 0.0f};
};
extern "C" freqdem freqdem_create(float kf) {
  if (kf <= 0.0f || !have_device()) return nullptr;
  freqdem q = new freqdem_s();
  q->kf = kf;
  q->ref = 1.0f / (2 * M_PI * kf);
  return q;
}
extern "C" int freqdem_demodulate_block(freqdem q, liquid_float_complex* r, unsigned int n, float* m) {
  if (!q || (!r && n) || (!m && n)) return LIQUID_EICONFIG;
  if (n == 0) return LIQUID_OK;
  Scratch& sc = scratch();
  if (sc.in.ensure((size_t)(n + 1) * 8) || sc.out.ensure((size_t)n * 4)) return LIQUID_EICONFIG;
  CUDA_TRY(cudaMemcpy(sc.in.p, &q->r_prime, 8, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy((float2*)sc.in.p + 1, r, (size_t)n * 8, cudaMemcpyHostToDevice));
  freqdem_kernel<<<(n + 255) / 256, 256>>>((const float2*)sc.in.p, (int)n, q->ref, (float*)sc.out.p);
  q->r_prime = ((const float2*)r)[n - 1];
  CUDA_TRY(cudaMemcpy(m, sc.out.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return LIQUID_OK;
}
extern "C" int freqdem_reset(freqdem q) {
  if (!q) return LIQUID_EICONFIG;
  q->r_prime = make_float2(0.0f, 0.0f);
  return LIQUID_OK;
}
extern "C" int freqdem_destroy(freqdem q) { delete q; return LIQUID_OK; }

// ============================================================================ firfilt
struct firfilt_rrrf_s {
  unsigned nt = 0;
  DevBuf taps, buf, out;     // buf = [history (nt - 1) | block]
  std::vector<float> hist;   // last nt - 1 inputs (host copy of the tail, re-uploaded with the next block)
};
extern "C" firfilt_rrrf firfilt_rrrf_create(float* h, unsigned int n) {
  if (!h || n == 0 || !have_device()) return nullptr;
  firfilt_rrrf q = new firfilt_rrrf_s();
  q->nt = n;
  q->hist.assign(n - 1, 0.0f);
  if (q->taps.alloc((size_t)n * 4)) { delete q; return nullptr; }
  cudaMemcpy(q->taps.p, h, (size_t)n * 4, cudaMemcpyHostToDevice);
  return q;
}
extern "C" int firfilt_rrrf_execute_block(firfilt_rrrf q, float* x, unsigned int n, float* y) {
  if (!q || (!x && n) || (!y && n)) return LIQUID_EICONFIG;
  if (n == 0) return LIQUID_OK;
  const unsigned nh = q->nt - 1;
  if (q->buf.ensure((size_t)(nh + n) * 4) || q->out.ensure((size_t)n * 4)) return LIQUID_EICONFIG;
  if (nh) CUDA_TRY(cudaMemcpy(q->buf.p, q->hist.data(), (size_t)nh * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy((float*)q->buf.p + nh, x, (size_t)n * 4, cudaMemcpyHostToDevice));
  // keep the tail before y is written: x == y is allowed (src/sdr_pmr446.c:896,901)
  if (nh) {
    std::vector<float> tail(nh);
    for (unsigned i = 0; i < nh; i++) {
      const long long src = (long long)n - nh + i;
      tail[i] = src >= 0 ? x[src] : q->hist[(size_t)(src + nh)];
    }
    q->hist.swap(tail);
  }
  fir_kernel<<<(n + 255) / 256, 256>>>((const float*)q->buf.p, (int)nh, (int)n, (const float*)q->taps.p, (int)q->nt, (float*)q->out.p);
  CUDA_TRY(cudaMemcpy(y, q->out.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return LIQUID_OK;
}
extern "C" int firfilt_rrrf_destroy(firfilt_rrrf q) { delete q; return LIQUID_OK; }

// ============================================================================ wdelayf, cbuffer (host containers, no arithmetic)
struct wdelayf_s {
  unsigned delay = 0, idx = 0;
  std::vector<float> v;
};
extern "C" wdelayf wdelayf_create(unsigned int delay) {
  wdelayf q = new wdelayf_s();
  q->delay = delay;
  q->v.assign(delay + 1, 0.0f);
  return q;
}
extern "C" int wdelayf_push(wdelayf q, float v) {
  q->v[q->idx] = v;
  q->idx = (q->idx + 1) % (q->delay + 1);
  return LIQUID_OK;
}
extern "C" int wdelayf_read(wdelayf q, float* v) {
  *v = q->v[q->idx];
  return LIQUID_OK;
}
extern "C" int wdelayf_destroy(wdelayf q) { delete q; return LIQUID_OK; }

template <typename T>
struct CBuf {
  std::vector<T> v;
  unsigned max = 0, num = 0, rd = 0;
};
struct cbuffercf_s { CBuf<liquid_float_complex> b; };
struct cbufferf_s { CBuf<float> b; };
template <typename T>
static int cb_write(CBuf<T>& q, const T* v, unsigned n) {
  if (n > q.max - q.num) return LIQUID_EIRANGE;
  if (q.rd + q.num + n > 2 * q.max) {
    memmove(q.v.data(), q.v.data() + q.rd, q.num * sizeof(T));
    q.rd = 0;
  }
  memcpy(q.v.data() + q.rd + q.num, v, n * sizeof(T));
  q.num += n;
  return LIQUID_OK;
}
template <typename T>
static int cb_release(CBuf<T>& q, unsigned n) {
  if (n > q.num) return LIQUID_EIRANGE;
  q.rd += n;
  q.num -= n;
  if (q.num == 0) q.rd = 0;
  return LIQUID_OK;
}
extern "C" cbuffercf cbuffercf_create(unsigned int max_size) {
  if (!max_size) return nullptr;
  cbuffercf q = new cbuffercf_s();
  q->b.max = max_size;
  q->b.v.resize(2 * (size_t)max_size);
  return q;
}
extern "C" int cbuffercf_write(cbuffercf q, liquid_float_complex* v, unsigned int n) { return cb_write(q->b, v, n); }
extern "C" unsigned int cbuffercf_size(cbuffercf q) { return q->b.num; }
extern "C" int cbuffercf_read(cbuffercf q, unsigned int n, liquid_float_complex** v, unsigned int* nr) {
  *nr = n > q->b.num ? q->b.num : n;
  *v = q->b.v.data() + q->b.rd;
  return LIQUID_OK;
}
extern "C" int cbuffercf_release(cbuffercf q, unsigned int n) { return cb_release(q->b, n); }
extern "C" int cbuffercf_destroy(cbuffercf q) { delete q; return LIQUID_OK; }
extern "C" cbufferf cbufferf_create(unsigned int max_size) {
  if (!max_size) return nullptr;
  cbufferf q = new cbufferf_s();
  q->b.max = max_size;
  q->b.v.resize(2 * (size_t)max_size);
  return q;
}
extern "C" int cbufferf_write(cbufferf q, float* v, unsigned int n) { return cb_write(q->b, v, n); }
extern "C" unsigned int cbufferf_size(cbufferf q) { return q->b.num; }
extern "C" unsigned int cbufferf_max_size(cbufferf q) { return q->b.max; }
extern "C" int cbufferf_read(cbufferf q, unsigned int n, float** v, unsigned int* nr) {
  *nr = n > q->b.num ? q->b.num : n;
  *v = q->b.v.data() + q->b.rd;
  return LIQUID_OK;
}
extern "C" int cbufferf_release(cbufferf q, unsigned int n) { return cb_release(q->b, n); }
extern "C" int cbufferf_destroy(cbufferf q) { delete q; return LIQUID_OK; }

// ============================================================================ asgram
struct asgramcf_s {
  Waterfall wf;
  unsigned W = 0;
  DevBuf ring, d_ascii, d_peak;
  long long cap = 0, n = 0;
};
extern "C" asgramcf asgramcf_create(unsigned int nfft) {
  if (nfft < 2 || !have_device()) return nullptr;
  asgramcf q = new asgramcf_s();
  q->W = nfft;
  q->cap = 1 << 18;
  if (q->wf.init(1, nfft) || q->ring.alloc_zero((size_t)q->cap * 8) || q->d_ascii.alloc(nfft) || q->d_peak.alloc(8)) { delete q; return nullptr; }
  q->wf.ref = 0.0f;      // asgramcf_create default scale; the reference then calls set_scale(-40, 2)
  q->wf.div = 10.0f;
  return q;
}
extern "C" int asgramcf_set_scale(asgramcf q, float ref, float div) {
  if (!q || div <= 0.0f) return LIQUID_EICONFIG;
  q->wf.ref = ref;
  q->wf.div = div;
  return LIQUID_OK;
}
extern "C" int asgramcf_write(asgramcf q, liquid_float_complex* x, unsigned int n) {
  if (!q || (!x && n)) return LIQUID_EICONFIG;
  if (q->n + n > q->cap) return LIQUID_EIRANGE;   // more than 2^18 samples between two execute() calls
  if (n) CUDA_TRY(cudaMemcpy((float2*)q->ring.p + q->n, x, (size_t)n * 8, cudaMemcpyHostToDevice));
  q->n += n;
  return LIQUID_OK;
}
extern "C" int asgramcf_execute(asgramcf q, char* ascii, float* peakval, float* peakfreq) {
  if (!q || !ascii || !peakval || !peakfreq) return LIQUID_EICONFIG;
  int launches = 0;
  if (q->wf.execute((const float2*)q->ring.p, q->cap, 0, q->n, (char*)q->d_ascii.p, (float*)q->d_peak.p, nullptr, nullptr, &launches))
    return LIQUID_EICONFIG;
  float pk[2];
  CUDA_TRY(cudaMemcpy(ascii, q->d_ascii.p, q->W, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(pk, q->d_peak.p, 8, cudaMemcpyDeviceToHost));
  *peakval = pk[0];
  *peakfreq = pk[1];
  q->n = 0;   // asgramcf_execute resets the periodogram
  return LIQUID_OK;
}
extern "C" int asgramcf_destroy(asgramcf q) {
  if (q) cudaDeviceSynchronize();
  delete q;
  return LIQUID_OK;
}
