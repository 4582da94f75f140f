/*
 * oracle/liquid_subset.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see header).
 *
 * Plain-C float32 restatement of the liquid-dsp v1.7.0 objects the reference calls
 * (call sites: /root/reference/src/sdr_pmr446.c:420-518,788-931; src/dsd_in.c:95-124,159-180).
 * Each block cites the SURVEY.md Appendix-A item it restates.  Compile with
 * -O2 -ffp-contract=off so results do not depend on the host's FMA support.
 */
#include "liquid_subset.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef float _Complex cf;

static oracle_liquid_knobs g_knobs = {256, 0, 0, 0, 0};
oracle_liquid_knobs *oracle_liquid_get_knobs(void) { return &g_knobs; }

/* ------------------------------------------------------------------ A.6 Kaiser design */
static float oracle_lngammaf(float z) { return lgammaf(z); }

float liquid_besseli0f(float z) {
  if (z == 0.0f) return 1.0f;
  float y = 0.0f;
  for (unsigned k = 0; k < 32; k++) {
    float t = (float)k * logf(0.5f * z) - oracle_lngammaf((float)k + 1.0f);
    y += expf(2.0f * t);
  }
  return y;
}

float kaiser_beta_As(float as) {
  as = fabsf(as);
  if (as > 50.0f) return 0.1102f * (as - 8.7f);
  if (as > 21.0f) return 0.5842f * powf(as - 21.0f, 0.4f) + 0.07886f * (as - 21.0f);
  return 0.0f;
}

float liquid_kaiser(unsigned i, unsigned wlen, float beta) {
  float t = (float)i - (float)(wlen - 1) / 2.0f;
  float r = (g_knobs.kaiser_r_mode == 0) ? 2.0f * t / (float)(wlen - 1) : 2.0f * t / (float)wlen;
  float a = liquid_besseli0f(beta * sqrtf(1.0f - r * r));
  float b = liquid_besseli0f(beta);
  return a / b;
}

static float oracle_sincf(float x) {
  if (fabsf(x) < 0.01f) return cosf((float)M_PI * x / 2.0f) * cosf((float)M_PI * x / 4.0f) * cosf((float)M_PI * x / 8.0f);
  return sinf((float)M_PI * x) / ((float)M_PI * x);
}

int liquid_firdes_kaiser(unsigned n, float fc, float as, float mu, float *h) {
  if (n == 0 || fc <= 0.0f || fc > 0.5f || mu < -0.5f || mu > 0.5f) return LIQUID_EICONFIG;
  float beta = kaiser_beta_As(as);
  for (unsigned i = 0; i < n; i++) {
    float t = (float)i - (float)(n - 1) / 2.0f + mu;
    float h1 = oracle_sincf(2.0f * fc * t);
    float h2 = liquid_kaiser(i, n, beta);
    h[i] = h1 * h2;
  }
  return LIQUID_OK;
}

unsigned estimate_req_filter_len(float df, float as) { return (unsigned)((as - 7.95f) / (14.26f * df)); }

/* ------------------------------------------------------------------ A.14 FFT (generic mixed radix) */
static void fft_rec(unsigned N, const cf *tw, unsigned n, unsigned istride, const cf *in, cf *out) {
  if (n == 1) {
    out[0] = in[0];
    return;
  }
  unsigned p = 2;
  while (n % p) p++;
  unsigned m = n / p;
  for (unsigned r = 0; r < p; r++) fft_rec(N, tw, m, istride * p, in + (size_t)r * istride, out + (size_t)r * m);
  unsigned tws = N / n;
  cf tmp[64];
  cf *t = (p <= 64) ? tmp : (cf *)malloc(p * sizeof(cf));
  for (unsigned k = 0; k < m; k++) {
    for (unsigned r = 0; r < p; r++) {
      cf w = tw[(size_t)((r * k) % n) * tws];
      cf v = out[r * m + k];
      t[r] = (crealf(v) * crealf(w) - cimagf(v) * cimagf(w)) + _Complex_I * (crealf(v) * cimagf(w) + cimagf(v) * crealf(w));
    }
    for (unsigned q = 0; q < p; q++) {
      float sr = 0.0f, si = 0.0f;
      for (unsigned r = 0; r < p; r++) {
        cf w = tw[(size_t)(((size_t)r * q * m) % n) * tws];
        sr += crealf(t[r]) * crealf(w) - cimagf(t[r]) * cimagf(w);
        si += crealf(t[r]) * cimagf(w) + cimagf(t[r]) * crealf(w);
      }
      out[q * m + k] = sr + _Complex_I * si;
    }
  }
  if (t != tmp) free(t);
}

typedef struct {
  unsigned n;
  cf *tw;
} fft_plan;
static fft_plan g_plans[16];
static unsigned g_nplans;
static const cf *fft_twiddles(unsigned n) {
  for (unsigned i = 0; i < g_nplans; i++)
    if (g_plans[i].n == n) return g_plans[i].tw;
  cf *tw = (cf *)malloc(n * sizeof(cf));
  for (unsigned k = 0; k < n; k++) {
    double a = -2.0 * M_PI * (double)k / (double)n;
    tw[k] = (float)cos(a) + _Complex_I * (float)sin(a);
  }
  if (g_nplans < 16) {
    g_plans[g_nplans].n = n;
    g_plans[g_nplans].tw = tw;
    g_nplans++;
  }
  return tw;
}
void oracle_fft_forward(unsigned n, const cf *in, cf *out) { fft_rec(n, fft_twiddles(n), n, 1, in, out); }

/* ------------------------------------------------------------------ A.1 iirfilt (Direct Form II) */
struct iirfilt_crcf_s {
  unsigned n;
  float b[4], a[4];
  cf v[4];
};
struct iirfilt_rrrf_s {
  unsigned n;
  float b[4], a[4];
  float v[4];
};

static int iir_setup(float *bo, float *ao, unsigned *n, const float *b, unsigned nb, const float *a, unsigned na) {
  if (nb == 0 || na == 0 || nb > 4 || na > 4) return -1;
  *n = nb > na ? nb : na;
  for (unsigned i = 0; i < 4; i++) bo[i] = ao[i] = 0.0f;
  float a0 = a[0];
  for (unsigned i = 0; i < nb; i++) bo[i] = b[i] / a0;
  for (unsigned i = 0; i < na; i++) ao[i] = a[i] / a0;
  return 0;
}
iirfilt_crcf iirfilt_crcf_create_dc_blocker(float alpha) {
  iirfilt_crcf q = (iirfilt_crcf)calloc(1, sizeof(*q));
  float b[2] = {1.0f, -1.0f}, a[2] = {1.0f, -1.0f + alpha};
  iir_setup(q->b, q->a, &q->n, b, 2, a, 2);
  return q;
}
int iirfilt_crcf_execute_block(iirfilt_crcf q, cf *x, unsigned n, cf *y) {
  for (unsigned i = 0; i < n; i++) {
    for (unsigned k = q->n - 1; k > 0; k--) q->v[k] = q->v[k - 1];
    cf v0 = 0;
    for (unsigned k = 1; k < q->n; k++) v0 += q->a[k] * q->v[k];
    v0 = x[i] - v0;
    q->v[0] = v0;
    cf acc = 0;
    for (unsigned k = 0; k < q->n; k++) acc += q->b[k] * q->v[k];
    y[i] = acc;
  }
  return LIQUID_OK;
}
int iirfilt_crcf_destroy(iirfilt_crcf q) {
  free(q);
  return LIQUID_OK;
}
iirfilt_rrrf iirfilt_rrrf_create(float *b, unsigned nb, float *a, unsigned na) {
  iirfilt_rrrf q = (iirfilt_rrrf)calloc(1, sizeof(*q));
  if (iir_setup(q->b, q->a, &q->n, b, nb, a, na)) {
    free(q);
    return NULL;
  }
  return q;
}
iirfilt_rrrf iirfilt_rrrf_create_dc_blocker(float alpha) {
  float b[2] = {1.0f, -1.0f}, a[2] = {1.0f, -1.0f + alpha};
  return iirfilt_rrrf_create(b, 2, a, 2);
}
int iirfilt_rrrf_execute_block(iirfilt_rrrf q, float *x, unsigned n, float *y) {
  for (unsigned i = 0; i < n; i++) {
    for (unsigned k = q->n - 1; k > 0; k--) q->v[k] = q->v[k - 1];
    float v0 = 0;
    for (unsigned k = 1; k < q->n; k++) v0 += q->a[k] * q->v[k];
    v0 = x[i] - v0;
    q->v[0] = v0;
    float acc = 0;
    for (unsigned k = 0; k < q->n; k++) acc += q->b[k] * q->v[k];
    y[i] = acc;
  }
  return LIQUID_OK;
}
int iirfilt_rrrf_destroy(iirfilt_rrrf q) {
  free(q);
  return LIQUID_OK;
}

/* ------------------------------------------------------------------ A.2-A.5 msresamp, both flavours */
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

#define T cf
#define X(name) CAT(name, _crcf)
#define API(fn) CAT(msresamp_crcf_, fn)
#include "resamp_tmpl.h"
#undef T
#undef X
#undef API

#define T float
#define X(name) CAT(name, _rrrf)
#define API(fn) CAT(msresamp_rrrf_, fn)
#include "resamp_tmpl.h"
#undef T
#undef X
#undef API

int oracle_msresamp_crcf_plan(msresamp_crcf q, unsigned *stages, unsigned *m_stage, float *rate_arb, unsigned *step, unsigned *npfb) {
  *stages = q->stages;
  for (unsigned i = 0; i < q->stages; i++) m_stage[i] = q->hb.m_stage[i];
  *rate_arb = q->rate_arb;
  *step = q->arb.step;
  *npfb = q->arb.npfb;
  return LIQUID_OK;
}

/* ------------------------------------------------------------------ A.7 nco (VCO) */
struct nco_crcf_s {
  unsigned theta, d_theta;
};
nco_crcf nco_crcf_create(liquid_ncotype type) {
  (void)type;
  return (nco_crcf)calloc(1, sizeof(struct nco_crcf_s));
}
int nco_crcf_set_frequency(nco_crcf q, float dtheta) {
  float p = dtheta * 0.159154943091895; /* float * double, rounded back to float */
  float fpart = p - ((long)p);
  if (fpart < 0.) fpart += 1.;
  q->d_theta = (unsigned)(fpart * 0xffffffff);
  return LIQUID_OK;
}
unsigned oracle_nco_crcf_get_dtheta_u32(nco_crcf q) { return q->d_theta; }
int nco_crcf_mix_down(nco_crcf q, cf x, cf *y) {
  float th = (float)(2.0 * M_PI) * ((float)q->theta / 4294967296.0f);
  float c = cosf(th), s = sinf(th);
  /* x * conj(c + js) */
  *y = (crealf(x) * c + cimagf(x) * s) + _Complex_I * (cimagf(x) * c - crealf(x) * s);
  return LIQUID_OK;
}
int nco_crcf_step(nco_crcf q) {
  q->theta += q->d_theta;
  return LIQUID_OK;
}
int nco_crcf_mix_block_down(nco_crcf q, cf *x, cf *y, unsigned n) {
  for (unsigned i = 0; i < n; i++) {
    nco_crcf_mix_down(q, x[i], &y[i]);
    nco_crcf_step(q);
  }
  return LIQUID_OK;
}
int nco_crcf_destroy(nco_crcf q) {
  free(q);
  return LIQUID_OK;
}

/* ------------------------------------------------------------------ A.8 firpfbch analyzer */
struct firpfbch_crcf_s {
  unsigned M, p;
  float *hsub; /* [M][p], oldest..newest */
  win_crcf *w;
  unsigned filter_index;
  cf *X, *x;
};
firpfbch_crcf firpfbch_crcf_create_kaiser(int type, unsigned M, unsigned m, float as) {
  if (type != LIQUID_ANALYZER || M == 0 || m == 0) return NULL;
  unsigned h_len = 2 * M * m + 1;
  float *h = (float *)malloc(h_len * sizeof(float));
  liquid_firdes_kaiser(h_len, 0.5f / (float)M, as, 0.0f, h);
  firpfbch_crcf q = (firpfbch_crcf)calloc(1, sizeof(*q));
  q->M = M;
  q->p = 2 * m;
  q->hsub = (float *)malloc((size_t)M * q->p * sizeof(float));
  q->w = (win_crcf *)calloc(M, sizeof(win_crcf));
  for (unsigned i = 0; i < M; i++) {
    for (unsigned n = 0; n < q->p; n++) q->hsub[i * q->p + (q->p - n - 1)] = h[i + n * M];
    win_init_crcf(&q->w[i], q->p);
  }
  free(h);
  q->filter_index = M - 1;
  q->X = (cf *)calloc(M, sizeof(cf));
  q->x = (cf *)calloc(M, sizeof(cf));
  return q;
}
int firpfbch_crcf_analyzer_execute(firpfbch_crcf q, cf *x, cf *y) {
  for (unsigned i = 0; i < q->M; i++) {
    win_push_crcf(&q->w[q->filter_index], x[i]);
    q->filter_index = (q->filter_index + q->M - 1) % q->M;
  }
  for (unsigned i = 0; i < q->M; i++) q->X[q->M - i - 1] = dot_crcf(q->hsub + (size_t)i * q->p, win_read_crcf(&q->w[i]), q->p);
  oracle_fft_forward(q->M, q->X, q->x);
  memmove(y, q->x, q->M * sizeof(cf));
  return LIQUID_OK;
}
int firpfbch_crcf_destroy(firpfbch_crcf q) {
  for (unsigned i = 0; i < q->M; i++) win_free_crcf(&q->w[i]);
  free(q->w);
  free(q->hsub);
  free(q->X);
  free(q->x);
  free(q);
  return LIQUID_OK;
}

/* ------------------------------------------------------------------ A.9 freqdem */
struct freqdem_s {
  float kf, ref;
  cf r_prime;
};
freqdem freqdem_create(float kf) {
  if (kf <= 0.0f) return NULL;
  freqdem q = (freqdem)calloc(1, sizeof(*q));
  q->kf = kf;
  q->ref = 1.0f / (2 * M_PI * kf);
  return q;
}
int freqdem_demodulate_block(freqdem q, cf *r, unsigned n, float *m) {
  for (unsigned i = 0; i < n; i++) {
    float pr = crealf(q->r_prime), pi = cimagf(q->r_prime);
    float re = pr * crealf(r[i]) + pi * cimagf(r[i]);
    float im = pr * cimagf(r[i]) - pi * crealf(r[i]);
    m[i] = atan2f(im, re) * q->ref;
    q->r_prime = r[i];
  }
  return LIQUID_OK;
}
int freqdem_reset(freqdem q) {
  q->r_prime = 0;
  return LIQUID_OK;
}
int freqdem_destroy(freqdem q) {
  free(q);
  return LIQUID_OK;
}

/* ------------------------------------------------------------------ A.10 firfilt */
struct firfilt_rrrf_s {
  unsigned n;
  float *hr; /* reversed: applied oldest..newest */
  win_rrrf w;
};
firfilt_rrrf firfilt_rrrf_create(float *h, unsigned n) {
  if (n == 0) return NULL;
  firfilt_rrrf q = (firfilt_rrrf)calloc(1, sizeof(*q));
  q->n = n;
  q->hr = (float *)malloc(n * sizeof(float));
  for (unsigned i = 0; i < n; i++) q->hr[n - 1 - i] = h[i];
  win_init_rrrf(&q->w, n);
  return q;
}
int firfilt_rrrf_execute_block(firfilt_rrrf q, float *x, unsigned n, float *y) {
  for (unsigned i = 0; i < n; i++) {
    win_push_rrrf(&q->w, x[i]);
    y[i] = dot_rrrf(q->hr, win_read_rrrf(&q->w), q->n);
  }
  return LIQUID_OK;
}
int firfilt_rrrf_destroy(firfilt_rrrf q) {
  win_free_rrrf(&q->w);
  free(q->hr);
  free(q);
  return LIQUID_OK;
}

/* ------------------------------------------------------------------ A.11 wdelay */
struct wdelayf_s {
  unsigned delay, idx;
  float *v;
};
wdelayf wdelayf_create(unsigned delay) {
  wdelayf q = (wdelayf)calloc(1, sizeof(*q));
  q->delay = delay;
  q->v = (float *)calloc(delay + 1, sizeof(float));
  return q;
}
int wdelayf_push(wdelayf q, float v) {
  q->v[q->idx] = v;
  q->idx = (q->idx + 1) % (q->delay + 1);
  return LIQUID_OK;
}
int wdelayf_read(wdelayf q, float *v) {
  *v = q->v[q->idx];
  return LIQUID_OK;
}
int wdelayf_destroy(wdelayf q) {
  free(q->v);
  free(q);
  return LIQUID_OK;
}

/* ------------------------------------------------------------------ A.12 cbuffer (linearised FIFO) */
#define CBUF_IMPL(NAME, TYPE)                                                                  \
  struct NAME##_s {                                                                            \
    TYPE *v;                                                                                   \
    unsigned max, num, rd;                                                                     \
  };                                                                                           \
  NAME NAME##_create(unsigned max_size) {                                                      \
    if (max_size == 0) return NULL;                                                            \
    NAME q = (NAME)calloc(1, sizeof(*q));                                                      \
    q->max = max_size;                                                                         \
    q->v = (TYPE *)calloc(2 * (size_t)max_size, sizeof(TYPE));                                 \
    return q;                                                                                  \
  }                                                                                            \
  int NAME##_write(NAME q, TYPE *v, unsigned n) {                                              \
    if (n > q->max - q->num) return LIQUID_EIRANGE;                                            \
    if (q->rd + q->num + n > 2 * q->max) {                                                     \
      memmove(q->v, q->v + q->rd, q->num * sizeof(TYPE));                                      \
      q->rd = 0;                                                                               \
    }                                                                                          \
    memcpy(q->v + q->rd + q->num, v, n * sizeof(TYPE));                                        \
    q->num += n;                                                                               \
    return LIQUID_OK;                                                                          \
  }                                                                                            \
  unsigned NAME##_size(NAME q) { return q->num; }                                              \
  int NAME##_read(NAME q, unsigned n, TYPE **v, unsigned *nr) {                                \
    *nr = n > q->num ? q->num : n;                                                             \
    *v = q->v + q->rd;                                                                         \
    return LIQUID_OK;                                                                          \
  }                                                                                            \
  int NAME##_release(NAME q, unsigned n) {                                                     \
    if (n > q->num) return LIQUID_EIRANGE;                                                     \
    q->rd += n;                                                                                \
    q->num -= n;                                                                               \
    if (q->num == 0) q->rd = 0;                                                                \
    return LIQUID_OK;                                                                          \
  }                                                                                            \
  int NAME##_destroy(NAME q) {                                                                 \
    free(q->v);                                                                                \
    free(q);                                                                                   \
    return LIQUID_OK;                                                                          \
  }
CBUF_IMPL(cbuffercf, cf)
CBUF_IMPL(cbufferf, float)
unsigned cbufferf_max_size(cbufferf q) { return q->max; }

/* ------------------------------------------------------------------ A.13 asgram over spgram */
struct asgramcf_s {
  unsigned nfft, p, nfftp;
  /* spgram */
  unsigned window_len, delay, sample_timer;
  unsigned long num_transforms;
  float *w;
  win_crcf buffer;
  cf *buf_time, *buf_freq;
  float *psd_acc;
  /* asgram */
  float *psd;
  float levels[10];
};
static const char asgram_levelchar[11] = " .,-+*&NM#";

static void spgram_clear(asgramcf q) {
  q->sample_timer = q->delay;
  q->num_transforms = 0;
  for (unsigned i = 0; i < q->nfftp; i++) q->psd_acc[i] = 0.0f;
}
asgramcf asgramcf_create(unsigned nfft) {
  if (nfft < 2) return NULL;
  asgramcf q = (asgramcf)calloc(1, sizeof(*q));
  q->nfft = nfft;
  q->p = 4;
  q->nfftp = nfft * q->p;
  q->window_len = nfft;
  q->delay = nfft / 2;
  q->w = (float *)malloc(nfft * sizeof(float));
  float g = 0.0f;
  for (unsigned i = 0; i < nfft; i++) {
    q->w[i] = 0.5f - 0.5f * cosf((2 * M_PI * (float)i) / ((float)(nfft - 1)));
    g += q->w[i] * q->w[i];
  }
  g = M_SQRT2 / (sqrtf(g / nfft) * sqrtf((float)(q->nfftp)));
  for (unsigned i = 0; i < nfft; i++) q->w[i] = g * q->w[i];
  win_init_crcf(&q->buffer, nfft);
  q->buf_time = (cf *)calloc(q->nfftp, sizeof(cf));
  q->buf_freq = (cf *)calloc(q->nfftp, sizeof(cf));
  q->psd_acc = (float *)calloc(q->nfftp, sizeof(float));
  q->psd = (float *)calloc(q->nfftp, sizeof(float));
  spgram_clear(q);
  asgramcf_set_scale(q, 0.0f, 10.0f);
  return q;
}
int asgramcf_set_scale(asgramcf q, float ref, float div) {
  if (div <= 0.0f) return LIQUID_EICONFIG;
  for (unsigned i = 0; i < 10; i++) q->levels[i] = ref + i * div;
  return LIQUID_OK;
}
int asgramcf_write(asgramcf q, cf *x, unsigned n) {
  for (unsigned s = 0; s < n; s++) {
    win_push_crcf(&q->buffer, x[s]);
    if (--q->sample_timer) continue;
    q->sample_timer = q->delay;
    const cf *rc = win_read_crcf(&q->buffer);
    for (unsigned i = 0; i < q->window_len; i++) q->buf_time[i] = rc[i] * q->w[i];
    oracle_fft_forward(q->nfftp, q->buf_time, q->buf_freq);
    for (unsigned i = 0; i < q->nfftp; i++) {
      float re = crealf(q->buf_freq[i]), im = cimagf(q->buf_freq[i]);
      float v = re * re + im * im;
      q->psd_acc[i] = (q->num_transforms == 0) ? v : q->psd_acc[i] + v;
    }
    q->num_transforms++;
  }
  return LIQUID_OK;
}
int asgramcf_execute(asgramcf q, char *ascii, float *peakval, float *peakfreq) {
  if (q->num_transforms == 0) {
    memset(ascii, ' ', q->nfft);
    *peakval = 0.0f;
    *peakfreq = 0.0f;
    return LIQUID_OK;
  }
  float scale = 1.0f / (float)(q->num_transforms > 1 ? q->num_transforms : 1);
  unsigned half = q->nfftp / 2;
  for (unsigned i = 0; i < q->nfftp; i++) {
    unsigned k = (i + half) % q->nfftp;
    float v = q->psd_acc[k] > 1e-12f ? q->psd_acc[k] : 1e-12f;
    q->psd[i] = 10 * log10f(v * scale);
  }
  spgram_clear(q);
  if (!g_knobs.asgram_keep_buf) {
    memset(q->buffer.buf, 0, 2 * (size_t)q->buffer.len * sizeof(cf));
    q->buffer.pos = 0;
  }
  for (unsigned i = 0; i < q->nfftp; i++) {
    if (i == 0 || q->psd[i] > *peakval) {
      *peakval = q->psd[i];
      *peakfreq = (float)i / (float)(q->nfftp) - 0.5f;
    }
  }
  for (unsigned i = 0; i < q->nfft; i++) {
    float val = 0.0f;
    for (unsigned j = 0; j < q->p; j++) {
      float v = q->psd[q->p * i + j];
      if (g_knobs.asgram_avg) val += v / (float)q->p;
      else val = (j == 0 || v > val) ? v : val;
    }
    ascii[i] = asgram_levelchar[0];
    for (unsigned j = 0; j < 10; j++)
      if (val > q->levels[j]) ascii[i] = asgram_levelchar[j];
  }
  return LIQUID_OK;
}
const float *oracle_asgramcf_last_psd(asgramcf q, unsigned *n) {
  *n = q->nfftp;
  return q->psd;
}
int asgramcf_destroy(asgramcf q) {
  free(q->w);
  win_free_crcf(&q->buffer);
  free(q->buf_time);
  free(q->buf_freq);
  free(q->psd_acc);
  free(q->psd);
  free(q);
  return LIQUID_OK;
}
