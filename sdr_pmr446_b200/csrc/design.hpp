// Host-side filter design for the B200 PMR446 chain.
//
// The reference builds every filter at start-up by calling liquid-dsp v1.7.0 constructors
// (/root/reference/src/sdr_pmr446.c:420-480, /root/reference/src/dsd_in.c:95-112):
//   msresamp_crcf_create(rate, 60)            -> half-band cascade + 14-tap arbitrary resampler
//   firpfbch_crcf_create_kaiser(ANALYZER,16,13,80) -> 416-tap polyphase prototype
//   asgramcf_create(W)                        -> scaled Hann window
// This file computes the same coefficient sets (float32, same formulas; SURVEY.md Appendix
// A.2-A.8, A.13) so the device kernels can run them.  It is product code: it does not use
// anything from oracle/.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace pmr {
namespace design {

// A.6 ---------------------------------------------------------------------------------------
inline float besseli0f(float z) {
  if (z == 0.0f) return 1.0f;
  float y = 0.0f;
  for (unsigned k = 0; k < 32; k++) {
    float t = (float)k * logf(0.5f * z) - lgammaf((float)k + 1.0f);
    y += expf(2.0f * t);
  }
  return y;
}
inline float kaiser_beta(float as) {
  as = fabsf(as);
  if (as > 50.0f) return 0.1102f * (as - 8.7f);
  if (as > 21.0f) return 0.5842f * powf(as - 21.0f, 0.4f) + 0.07886f * (as - 21.0f);
  return 0.0f;
}
inline float kaiser_win(unsigned i, unsigned n, float beta) {
  float t = (float)i - (float)(n - 1) / 2.0f;
  float r = 2.0f * t / (float)(n - 1);
  return besseli0f(beta * sqrtf(1.0f - r * r)) / besseli0f(beta);
}
inline float sincf_(float x) {
  const float pi = (float)M_PI;
  if (fabsf(x) < 0.01f) return cosf(pi * x / 2.0f) * cosf(pi * x / 4.0f) * cosf(pi * x / 8.0f);
  return sinf(pi * x) / (pi * x);
}
inline std::vector<float> firdes_kaiser(unsigned n, float fc, float as, float mu = 0.0f) {
  std::vector<float> h(n);
  float beta = kaiser_beta(as);
  for (unsigned i = 0; i < n; i++) {
    float t = (float)i - (float)(n - 1) / 2.0f + mu;
    h[i] = sincf_(2.0f * fc * t) * kaiser_win(i, n, beta);
  }
  return h;
}

// A.4: one half-band stage of semi-length m.  Returns the 2m filter-branch taps in
// "newest first" order: out = x_odd[o - m] + sum_j taps[j] * x_even[o - j].
inline std::vector<float> halfband_taps(unsigned m, float as) {
  unsigned n = 4 * m + 1;
  float beta = kaiser_beta(as);
  std::vector<float> h(n);
  for (unsigned i = 0; i < n; i++) {
    float t = (float)i - (float)(n - 1) / 2.0f;
    h[i] = sincf_(t / 2.0f) * kaiser_win(i, n, beta);
  }
  std::vector<float> taps(2 * m);
  for (unsigned j = 0; j < 2 * m; j++) taps[j] = h[2 * j + 1];
  return taps;
}

// A.2/A.3/A.5: the multi-stage resampler plan.
struct MsresampPlan {
  float rate = 1.0f, as = 60.0f;
  bool interp = false;
  unsigned stages = 0;            // number of half-band stages
  std::vector<unsigned> m;        // [stages]; index 0 runs at the LOWEST rate
  std::vector<std::vector<float>> hb;  // hb[g] = halfband_taps(m[g])
  float zeta = 1.0f;              // 2^-stages, applied once after the decimating cascade
  float rate_arb = 1.0f;
  unsigned arb_m = 7, npfb = 256, bits = 8, sub_len = 14;
  uint32_t step = 1u << 24;       // 24-bit fixed-point phase increment
  std::vector<float> pfb;         // [npfb][sub_len], newest-first: y = sum_k pfb[idx][k] * u[i-k]
};

inline MsresampPlan msresamp_plan(float rate, float as, unsigned npfb_req = 256) {
  MsresampPlan p;
  p.rate = rate;
  p.as = as;
  p.interp = rate > 1.0f;
  p.rate_arb = rate;
  if (p.interp)
    while (p.rate_arb > 2.0f) { p.stages++; p.rate_arb *= 0.5f; }
  else
    while (p.rate_arb < 0.5f) { p.stages++; p.rate_arb *= 2.0f; }
  float fc = 0.4f, as5 = as + 5.0f;
  for (unsigned i = 0; i < p.stages; i++) {
    fc = (i == 1) ? (0.5f - fc) / 2.0f : 0.5f * fc;
    float ft = 2.0f * (0.25f - fc);
    unsigned h_len = (unsigned)((as5 - 7.95f) / (14.26f * ft));
    unsigned m = (unsigned)ceilf((float)(h_len - 1) / 4.0f);
    if (m < 3) m = 3;
    p.m.push_back(m);
    p.hb.push_back(halfband_taps(m, as5));
  }
  p.zeta = 1.0f / (float)(1u << p.stages);
  p.bits = 0;
  while ((1u << p.bits) < npfb_req) p.bits++;
  p.npfb = 1u << p.bits;
  unsigned n = 2 * p.arb_m * p.npfb + 1;
  float fca = fminf(0.49f, 0.515f * p.rate_arb);
  std::vector<float> hf = firdes_kaiser(n, fca / (float)p.npfb, as);
  float gain = 0.0f;
  for (unsigned i = 0; i < n; i++) gain += hf[i];
  gain = (float)p.npfb / gain;
  for (unsigned i = 0; i < n; i++) hf[i] *= gain;
  p.sub_len = (n - 1) / p.npfb;
  p.pfb.resize((size_t)p.npfb * p.sub_len);
  for (unsigned i = 0; i < p.npfb; i++)
    for (unsigned k = 0; k < p.sub_len; k++) p.pfb[(size_t)i * p.sub_len + k] = hf[i + k * p.npfb];
  p.step = (uint32_t)round((1 << 24) / p.rate_arb);
  return p;
}

// Number of arbitrary-resampler outputs emitted once `nu` inputs have been pushed (A.5):
// output j is emitted with input floor(j*step / 2^24), so count = ceil(nu * 2^24 / step).
inline uint64_t arb_outputs_after(uint64_t nu, uint32_t step) {
  if (nu == 0) return 0;
  unsigned __int128 num = (unsigned __int128)nu << 24;
  return (uint64_t)((num + step - 1) / step);
}

// A.8: analysis channelizer prototype.  taps[i*p + n] = h[i + n*M], i < M, n < p = 2m (newest first).
inline std::vector<float> pfbch_taps(unsigned M, unsigned m, float as) {
  unsigned h_len = 2 * M * m + 1;
  std::vector<float> h = firdes_kaiser(h_len, 0.5f / (float)M, as);
  unsigned p = 2 * m;
  std::vector<float> t((size_t)M * p);
  for (unsigned i = 0; i < M; i++)
    for (unsigned n = 0; n < p; n++) t[(size_t)i * p + n] = h[i + n * M];
  return t;
}

// A.7: NCO frequency word.  Same float32 steps as nco_crcf_set_frequency().
inline uint32_t nco_dtheta(float dtheta) {
  float p = dtheta * 0.159154943091895;
  float fpart = p - ((long)p);
  if (fpart < 0.) fpart += 1.;
  return (uint32_t)(fpart * 0xffffffff);
}

// A.13: spgram window (Hann of length W scaled for an nfft = 4W transform).
inline std::vector<float> asgram_window(unsigned W) {
  std::vector<float> w(W);
  float g = 0.0f;
  for (unsigned i = 0; i < W; i++) {
    w[i] = 0.5f - 0.5f * cosf((2 * M_PI * (float)i) / ((float)(W - 1)));
    g += w[i] * w[i];
  }
  g = M_SQRT2 / (sqrtf(g / W) * sqrtf((float)(4 * W)));
  for (unsigned i = 0; i < W; i++) w[i] = g * w[i];
  return w;
}

}  // namespace design
}  // namespace pmr
