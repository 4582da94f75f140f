/*
 * oracle/resamp_tmpl.h -- TEST INFRASTRUCTURE.  Included twice by liquid_subset.c with
 *   T   = sample type (float _Complex | float),  X(name) = name##_crcf | name##_rrrf
 * Restates liquid-dsp v1.7.0 msresamp / msresamp2 / resamp2 / resamp (fixed-point phase) /
 * firpfb as used by the reference at /root/reference/src/sdr_pmr446.c:425-426,796 and
 * /root/reference/src/dsd_in.c:100,104,168,170.
 * Coefficients are real in both instantiations.
 */

/* ---- sliding window: read pointer is oldest..newest, contiguous (doubled storage) ---- */
typedef struct {
  T *buf;
  unsigned len, pos;
} X(win);

static void X(win_init)(X(win) * w, unsigned len) {
  w->len = len;
  w->pos = 0;
  w->buf = (T *)calloc(2 * (size_t)len, sizeof(T));
}
static void X(win_free)(X(win) * w) { free(w->buf); }
static inline void X(win_push)(X(win) * w, T v) {
  w->buf[w->pos] = v;
  w->buf[w->pos + w->len] = v;
  if (++w->pos == w->len) w->pos = 0;
}
static inline const T *X(win_read)(const X(win) * w) { return w->buf + w->pos; }

/* dot product, real taps, single accumulator in index order (liquid portable-C dotprod, A.14) */
static inline T X(dot)(const float *h, const T *x, unsigned n) {
  T r = 0;
  for (unsigned i = 0; i < n; i++) r += h[i] * x[i];
  return r;
}

/* ---- resamp2: one half-band stage (A.4) ---- */
typedef struct {
  unsigned m;
  float *h1;   /* 2m taps: h1[j] = h[4m-1-2j], applied oldest..newest */
  X(win) w0, w1;
} X(resamp2);

static void X(resamp2_init)(X(resamp2) * q, unsigned m, float as) {
  q->m = m;
  unsigned h_len = 4 * m + 1;
  float *h = (float *)malloc(h_len * sizeof(float));
  float beta = kaiser_beta_As(as);
  for (unsigned i = 0; i < h_len; i++) {
    float t = (float)i - (float)(h_len - 1) / 2.0f;
    float h1 = oracle_sincf(t / 2.0f);
    float h2 = liquid_kaiser(i, h_len, beta);
    h[i] = h1 * h2; /* f0 = 0 -> modulation term cos(0) = 1 */
  }
  q->h1 = (float *)malloc(2 * m * sizeof(float));
  unsigned j = 0;
  for (unsigned i = 1; i < h_len; i += 2) q->h1[j++] = h[h_len - i - 1];
  free(h);
  X(win_init)(&q->w0, 2 * m);
  X(win_init)(&q->w1, 2 * m);
}
static void X(resamp2_free)(X(resamp2) * q) {
  free(q->h1);
  X(win_free)(&q->w0);
  X(win_free)(&q->w1);
}
/* decimator: x[0] feeds the filter branch, x[1] the delay branch; DC gain 2 */
static inline T X(resamp2_decim)(X(resamp2) * q, const T *x) {
  X(win_push)(&q->w1, x[0]);
  T y1 = X(dot)(q->h1, X(win_read)(&q->w1), 2 * q->m);
  X(win_push)(&q->w0, x[1]);
  T y0 = X(win_read)(&q->w0)[q->m - 1];
  return y0 + y1;
}
/* interpolator: y[0] = delay branch, y[1] = filter branch */
static inline void X(resamp2_interp)(X(resamp2) * q, T x, T *y) {
  X(win_push)(&q->w0, x);
  y[0] = X(win_read)(&q->w0)[q->m - 1];
  X(win_push)(&q->w1, x);
  y[1] = X(dot)(q->h1, X(win_read)(&q->w1), 2 * q->m);
}

/* ---- msresamp2: half-band cascade (A.3) ---- */
typedef struct {
  int interp;
  unsigned stages;
  unsigned *m_stage;
  X(resamp2) * s;
  T *b0, *b1;
  float zeta;
} X(msresamp2);

static void X(msresamp2_init)(X(msresamp2) * q, int interp, unsigned stages, float fc, float as_in) {
  q->interp = interp;
  q->stages = stages;
  q->m_stage = (unsigned *)calloc(stages ? stages : 1, sizeof(unsigned));
  q->s = (X(resamp2) *)calloc(stages ? stages : 1, sizeof(X(resamp2)));
  q->b0 = (T *)calloc((size_t)1 << stages, sizeof(T));
  q->b1 = (T *)calloc((size_t)1 << stages, sizeof(T));
  float as = as_in + 5.0f;
  for (unsigned i = 0; i < stages; i++) {
    fc = (i == 1) ? (0.5f - fc) / 2.0f : 0.5f * fc;
    float ft = 2.0f * (0.25f - fc);
    unsigned h_len = estimate_req_filter_len(ft, as);
    unsigned m = (unsigned)ceilf((float)(h_len - 1) / 4.0f);
    q->m_stage[i] = m < 3 ? 3 : m;
    X(resamp2_init)(&q->s[i], q->m_stage[i], as);
  }
  q->zeta = 1.0f / (float)(1u << stages);
}
static void X(msresamp2_free)(X(msresamp2) * q) {
  for (unsigned i = 0; i < q->stages; i++) X(resamp2_free)(&q->s[i]);
  free(q->s);
  free(q->m_stage);
  free(q->b0);
  free(q->b1);
}
/* 2^stages inputs -> 1 output; highest-index stage runs first (at the highest rate) */
static T X(msresamp2_decim)(X(msresamp2) * q, const T *x) {
  const T *in = x;
  T *out = q->b1;
  for (unsigned s = 0; s < q->stages; s++) {
    unsigned g = q->stages - s - 1;
    unsigned n = 1u << g;
    for (unsigned k = 0; k < n; k++) out[k] = X(resamp2_decim)(&q->s[g], &in[2 * k]);
    in = out;
    out = (out == q->b1) ? q->b0 : q->b1;
  }
  return in[0] * q->zeta;
}
/* 1 input -> 2^stages outputs; stage 0 first, no scaling */
static void X(msresamp2_interp)(X(msresamp2) * q, T x, T *y) {
  T *in = q->b0, *out = q->b1;
  in[0] = x;
  for (unsigned s = 0; s < q->stages; s++) {
    unsigned n = 1u << s;
    T *dst = (s == q->stages - 1) ? y : out;
    for (unsigned k = 0; k < n; k++) X(resamp2_interp)(&q->s[s], in[k], &dst[2 * k]);
    T *t = in;
    in = out;
    out = t;
  }
  if (q->stages == 0) y[0] = x;
}

/* ---- resamp: arbitrary-rate polyphase resampler with 24-bit fixed-point phase (A.5) ---- */
typedef struct {
  unsigned m, npfb, bits, sub_len;
  float *hsub; /* [npfb][sub_len], each stored oldest..newest (reverse of prototype order) */
  X(win) w;
  unsigned step, phase;
} X(resamp);

static void X(resamp_init)(X(resamp) * q, float rate, unsigned m, float fc, float as, unsigned npfb_req) {
  unsigned bits = 0;
  while ((1u << bits) < npfb_req) bits++;
  q->bits = bits;
  q->npfb = 1u << bits;
  q->m = m;
  unsigned n = 2 * m * q->npfb + 1;
  float *hf = (float *)malloc(n * sizeof(float));
  liquid_firdes_kaiser(n, fc / (float)q->npfb, as, 0.0f, hf);
  float gain = 0.0f;
  for (unsigned i = 0; i < n; i++) gain += hf[i];
  gain = (float)q->npfb / gain;
  for (unsigned i = 0; i < n; i++) hf[i] *= gain;
  q->sub_len = (n - 1) / q->npfb; /* last prototype tap dropped */
  q->hsub = (float *)malloc((size_t)q->npfb * q->sub_len * sizeof(float));
  for (unsigned i = 0; i < q->npfb; i++)
    for (unsigned k = 0; k < q->sub_len; k++) q->hsub[i * q->sub_len + (q->sub_len - k - 1)] = hf[i + k * q->npfb];
  free(hf);
  X(win_init)(&q->w, q->sub_len);
  q->step = (unsigned)round((1 << 24) / rate); /* int / float -> float32 division, as liquid does */
  q->phase = 0;
}
static void X(resamp_free)(X(resamp) * q) {
  free(q->hsub);
  X(win_free)(&q->w);
}
static inline unsigned X(resamp_exec)(X(resamp) * q, T x, T *y) {
  X(win_push)(&q->w, x);
  unsigned n = 0;
  while (q->phase < (1u << 24)) {
    unsigned idx = q->phase >> (24 - q->bits);
    y[n++] = X(dot)(q->hsub + (size_t)idx * q->sub_len, X(win_read)(&q->w), q->sub_len);
    q->phase += q->step;
  }
  q->phase -= (1u << 24);
  return n;
}

/* ---- msresamp (A.2) ---- */
struct X(msresamp_s) {
  float rate, as, rate_arb;
  int interp;
  unsigned stages;
  X(msresamp2) hb;
  X(resamp) arb;
  T *buffer;
  unsigned buffer_index;
};

X(msresamp) API(create)(float rate, float as) {
  if (!(rate > 0.0f)) return NULL;
  X(msresamp) q = (X(msresamp))calloc(1, sizeof(*q));
  q->rate = rate;
  q->as = as;
  q->interp = rate > 1.0f;
  q->rate_arb = rate;
  q->stages = 0;
  if (q->interp) {
    while (q->rate_arb > 2.0f) { q->stages++; q->rate_arb *= 0.5f; }
  } else {
    while (q->rate_arb < 0.5f) { q->stages++; q->rate_arb *= 2.0f; }
  }
  q->buffer = (T *)calloc(4 + ((size_t)1 << q->stages), sizeof(T));
  q->buffer_index = 0;
  X(msresamp2_init)(&q->hb, q->interp, q->stages, 0.4f, 0.0f + as);
  const oracle_liquid_knobs *kn = oracle_liquid_get_knobs();
  float fc = (kn->resamp_fc_mode == 0) ? fminf(0.49f, 0.515f * q->rate_arb) : 0.4f;
  X(resamp_init)(&q->arb, q->rate_arb, 7, fc, as, (unsigned)kn->resamp_npfb);
  return q;
}
int API(destroy)(X(msresamp) q) {
  if (!q) return LIQUID_EICONFIG;
  X(msresamp2_free)(&q->hb);
  X(resamp_free)(&q->arb);
  free(q->buffer);
  free(q);
  return LIQUID_OK;
}
int API(print)(X(msresamp) q) {
  printf("<oracle msresamp rate=%g, %s, halfband stages=%u (m:", q->rate, q->interp ? "interp" : "decim", q->stages);
  for (unsigned i = 0; i < q->stages; i++) printf(" %u", q->hb.m_stage[i]);
  printf("), arbitrary rate=%g step=%u npfb=%u>\n", q->rate_arb, q->arb.step, q->arb.npfb);
  return LIQUID_OK;
}
int API(execute)(X(msresamp) q, T *x, unsigned nx, T *y, unsigned *ny_out) {
  unsigned ny = 0;
  if (!q->interp) {
    unsigned M = 1u << q->stages;
    for (unsigned i = 0; i < nx; i++) {
      q->buffer[q->buffer_index++] = x[i];
      if (q->buffer_index == M) {
        T hb = X(msresamp2_decim)(&q->hb, q->buffer);
        ny += X(resamp_exec)(&q->arb, hb, &y[ny]);
        q->buffer_index = 0;
      }
    }
  } else {
    unsigned M = 1u << q->stages;
    for (unsigned i = 0; i < nx; i++) {
      unsigned nw = X(resamp_exec)(&q->arb, x[i], q->buffer);
      for (unsigned k = 0; k < nw; k++) {
        X(msresamp2_interp)(&q->hb, q->buffer[k], &y[ny]);
        ny += M;
      }
    }
  }
  *ny_out = ny;
  return LIQUID_OK;
}
