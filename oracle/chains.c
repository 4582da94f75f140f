/*
 * oracle/chains.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see liquid_subset.h).
 *
 * The two processing chains of the reference, restated over the liquid-subset oracle:
 *   - PMR446:  /root/reference/src/sdr_pmr446.c:420-480 (object parameters) and :788-913
 *              (order of operations, buffers, carry-over of <16 samples in the ring buffer).
 *   - dsd_in:  /root/reference/src/dsd_in.c:95-112 and :159-180.
 * Differences from the reference, all deliberate (SURVEY.md "read this first" #3, §8a):
 *   - the SoapySDR read is replaced by a caller-supplied buffer; cu8 input is converted with
 *     the SoapyRTLSDR rule (u8 - 127.4)/128 (row a0);
 *   - the demodulation chain (:881-902) is instantiated for EVERY channel, each with its own
 *     filter state from t = 0, instead of only for the squelch-selected one; setting
 *     cfg.active_only >= 0 gives the reference-faithful single-channel cost for CPU timing;
 *   - the squelch state machine, CTCSS detector, RtAudio and terminal output are out of scope;
 *   - sample rate, channel count and chunk size are configuration instead of #defines;
 *   - s16 output for the PMR chain uses dsd_in's conversion (src/dsd_in.c:172-175), saturated at +-full scale (the
 *     reference's PMR audio is float32 clipped by the audio device, never a wrapping cast).
 */
#include "chains.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pmr446_taps.h"

typedef float _Complex cf;

static inline int16_t to_s16(float v) { return (int16_t)(int32_t)(v * (float)INT16_MAX); }   /* dsd_in: plain C cast */
static inline int16_t to_s16_sat(float v) {
  float y = v * (float)INT16_MAX;
  y = y < -32768.0f ? -32768.0f : (y > 32767.0f ? 32767.0f : y);
  return (int16_t)(int32_t)y;
}

static void convert_in(int fmt, const void *iq, unsigned n, cf *out) {
  if (fmt == ORACLE_FMT_CF32) {
    memcpy(out, iq, (size_t)n * sizeof(cf));
  } else {
    const uint8_t *b = (const uint8_t *)iq;
    for (unsigned i = 0; i < n; i++) {
      float re = ((float)b[2 * i] - 127.4f) * (1.0f / 128.0f);
      float im = ((float)b[2 * i + 1] - 127.4f) * (1.0f / 128.0f);
      out[i] = re + _Complex_I * im;
    }
  }
}

/* ------------------------------------------------------------------ PMR446 chain */
struct oracle_pmr_s {
  oracle_pmr_cfg cfg;
  unsigned M;
  iirfilt_crcf dcblock;
  msresamp_crcf resampler;
  nco_crcf nco;
  firpfbch_crcf channelizer;
  freqdem *fm_demod;        /* [M] */
  firfilt_rrrf *ctcss_filt; /* [M] */
  wdelayf *ctcss_lp_delay;  /* [M] */
  firfilt_rrrf *audio_filt; /* [M] */
  iirfilt_rrrf *deemph;     /* [M] */
  firfilt_rrrf *deemph_fir; /* [M], APP_FIR_DEEMPH variant */
  cbuffercf resamp_buf;
  asgramcf asgram;
  cf *buffp, *resamp_tmp, *chan_tmp;
  float *tmp1, *tmp2;
  unsigned res_size, chan_size;
};

void oracle_pmr_default_cfg(oracle_pmr_cfg *c) {
  memset(c, 0, sizeof(*c));
  c->fs_in = 1024000;  /* include/sdr_pmr446.h:13 */
  c->in_fmt = ORACLE_FMT_CF32;
  c->num_channels = 16;   /* src/sdr_pmr446.c:23 */
  c->channel_width = 12500; /* :22 */
  c->pfb_m = 13;          /* :437 */
  c->pfb_as = 80.0f;
  c->resamp_as = 60.0f;   /* :426 */
  c->dc_alpha = 0.0005f;  /* :422 */
  c->kf = 0.5f;           /* :440 */
  c->audio_gain = 4.0f;   /* :33 */
  c->lowpass = 0;
  c->waterfall = 0;
  c->chunk = 100000;      /* :30 */
  c->active_only = -1;
}

oracle_pmr *oracle_pmr_create(const oracle_pmr_cfg *cfg) {
  oracle_pmr *o = (oracle_pmr *)calloc(1, sizeof(*o));
  o->cfg = *cfg;
  unsigned M = o->M = cfg->num_channels;
  float fs_res = (float)(M * cfg->channel_width);
  o->dcblock = iirfilt_crcf_create_dc_blocker(cfg->dc_alpha);
  o->resampler = msresamp_crcf_create(fs_res / (float)cfg->fs_in, cfg->resamp_as);
  o->nco = nco_crcf_create(LIQUID_VCO);
  float offset = -0.5f * (float)(M - 1) / (float)M * 2 * M_PI; /* :432-433 */
  nco_crcf_set_frequency(o->nco, offset);
  o->channelizer = firpfbch_crcf_create_kaiser(LIQUID_ANALYZER, M, cfg->pfb_m, cfg->pfb_as);
  if (!o->dcblock || !o->resampler || !o->nco || !o->channelizer) return NULL;

  float hp[PMR446_HP_AUDIO_TAPS_LEN], lp[PMR446_LP_AUDIO_TAPS_LEN];
  pmr446_hp_audio_taps_fill(hp);
  pmr446_lp_audio_taps_fill(lp);
  o->fm_demod = (freqdem *)calloc(M, sizeof(freqdem));
  o->ctcss_filt = (firfilt_rrrf *)calloc(M, sizeof(firfilt_rrrf));
  o->ctcss_lp_delay = (wdelayf *)calloc(M, sizeof(wdelayf));
  o->audio_filt = (firfilt_rrrf *)calloc(M, sizeof(firfilt_rrrf));
  o->deemph = (iirfilt_rrrf *)calloc(M, sizeof(iirfilt_rrrf));
  o->deemph_fir = (firfilt_rrrf *)calloc(M, sizeof(firfilt_rrrf));
  float de[PMR446_FIR_DEEMPH_TAPS_LEN];
  pmr446_fir_deemph_taps_fill(de);
  for (unsigned i = 0; i < M; i++) {
    o->fm_demod[i] = freqdem_create(cfg->kf);
    o->ctcss_filt[i] = firfilt_rrrf_create(hp, PMR446_HP_AUDIO_TAPS_LEN);
    o->ctcss_lp_delay[i] = wdelayf_create((PMR446_HP_AUDIO_TAPS_LEN - 1) / 2);
    o->audio_filt[i] = firfilt_rrrf_create(lp, PMR446_LP_AUDIO_TAPS_LEN);
    o->deemph[i] = iirfilt_rrrf_create((float[]){PMR446_DEEMPH_B0, PMR446_DEEMPH_B1}, 2,
                                       (float[]){PMR446_DEEMPH_A0, PMR446_DEEMPH_A1}, 2);
    o->deemph_fir[i] = firfilt_rrrf_create(de, PMR446_FIR_DEEMPH_TAPS_LEN); /* :458 */
  }
  /* buffer sizes, :730-732 */
  o->res_size = (unsigned)ceilf(1 + 2 * cfg->chunk * (fs_res / (float)cfg->fs_in));
  o->chan_size = (unsigned)ceilf(o->res_size / M) + 1;
  o->resamp_buf = cbuffercf_create(o->res_size + M);
  if (cfg->waterfall > 0) {
    o->asgram = asgramcf_create(cfg->waterfall);
    asgramcf_set_scale(o->asgram, -40.0f, 2.0f); /* :476 */
  }
  o->buffp = (cf *)malloc((size_t)cfg->chunk * sizeof(cf));
  o->resamp_tmp = (cf *)malloc((size_t)o->res_size * sizeof(cf));
  o->chan_tmp = (cf *)malloc((size_t)M * o->chan_size * sizeof(cf));
  o->tmp1 = (float *)malloc((size_t)o->chan_size * sizeof(float));
  o->tmp2 = (float *)malloc((size_t)o->chan_size * sizeof(float));
  return o;
}

void oracle_pmr_destroy(oracle_pmr *o) {
  if (!o) return;
  for (unsigned i = 0; i < o->M; i++) {
    freqdem_destroy(o->fm_demod[i]);
    firfilt_rrrf_destroy(o->ctcss_filt[i]);
    wdelayf_destroy(o->ctcss_lp_delay[i]);
    firfilt_rrrf_destroy(o->audio_filt[i]);
    iirfilt_rrrf_destroy(o->deemph[i]);
    firfilt_rrrf_destroy(o->deemph_fir[i]);
  }
  free(o->deemph_fir);
  free(o->fm_demod); free(o->ctcss_filt); free(o->ctcss_lp_delay); free(o->audio_filt); free(o->deemph);
  if (o->asgram) asgramcf_destroy(o->asgram);
  cbuffercf_destroy(o->resamp_buf);
  firpfbch_crcf_destroy(o->channelizer);
  nco_crcf_destroy(o->nco);
  msresamp_crcf_destroy(o->resampler);
  iirfilt_crcf_destroy(o->dcblock);
  free(o->buffp); free(o->resamp_tmp); free(o->chan_tmp); free(o->tmp1); free(o->tmp2);
  free(o);
}

unsigned oracle_pmr_res_size(const oracle_pmr *o) { return o->res_size; }
unsigned oracle_pmr_chan_size(const oracle_pmr *o) { return o->chan_size; }

int oracle_pmr_execute(oracle_pmr *o, const void *iq, unsigned n, const oracle_pmr_out *out, unsigned *ny_out, unsigned *ns_out) {
  const unsigned M = o->M;
  if (n > o->cfg.chunk) return -1;
  cf *buffp = o->buffp;
  convert_in(o->cfg.in_fmt, iq, n, buffp);

  /* :795-797 */
  unsigned ny = 0;
  iirfilt_crcf_execute_block(o->dcblock, buffp, n, buffp);
  if (out->dcblocked) memcpy(out->dcblocked, buffp, (size_t)n * sizeof(cf));
  msresamp_crcf_execute(o->resampler, buffp, n, o->resamp_tmp, &ny);
  if (out->res) memcpy(out->res, o->resamp_tmp, (size_t)ny * sizeof(cf));
  if (cbuffercf_write(o->resamp_buf, o->resamp_tmp, ny) != LIQUID_OK) return -2;

  /* :804-823 */
  unsigned ns = 0, num_read;
  cf *rpc;
  cf *tmp_out = (cf *)alloca(M * sizeof(cf));
  while (cbuffercf_size(o->resamp_buf) >= M) {
    cbuffercf_read(o->resamp_buf, M, &rpc, &num_read);
    for (unsigned i = 0; i < M; i++) {
      nco_crcf_mix_down(o->nco, rpc[i], &rpc[i]);
      nco_crcf_step(o->nco);
    }
    firpfbch_crcf_analyzer_execute(o->channelizer, rpc, tmp_out);
    cbuffercf_release(o->resamp_buf, num_read);
    for (unsigned i = 0; i < M; i++) o->chan_tmp[(size_t)i * o->chan_size + ns] = tmp_out[i];
    ns++;
  }
  if (ns > o->chan_size || (out->ld && ns > out->ld)) return -3;

  /* :876-908, for every channel (or only cfg.active_only) */
  for (unsigned i = 0; i < M; i++) {
    if (o->cfg.active_only >= 0 && (unsigned)o->cfg.active_only != i) continue;
    cf *cb = o->chan_tmp + (size_t)i * o->chan_size;
    float *t1 = o->tmp1, *t2 = o->tmp2;
    if (out->chan) memcpy(out->chan + (size_t)i * out->ld, cb, (size_t)ns * sizeof(cf));
    if (o->cfg.channelize_only) continue;
    freqdem_demodulate_block(o->fm_demod[i], cb, ns, t1);
    if (out->demod) memcpy(out->demod + (size_t)i * out->ld, t1, (size_t)ns * sizeof(float));
    firfilt_rrrf_execute_block(o->ctcss_filt[i], t1, ns, t2);
    for (unsigned k = 0; k < ns; k++) {
      float tmp;
      wdelayf_push(o->ctcss_lp_delay[i], t1[k]);
      wdelayf_read(o->ctcss_lp_delay[i], &tmp);
      t1[k] = tmp - t2[k];
      t2[k] *= o->cfg.audio_gain;
    }
    if (out->lpcomp) memcpy(out->lpcomp + (size_t)i * out->ld, t1, (size_t)ns * sizeof(float));
    if (o->cfg.deemph_fir) firfilt_rrrf_execute_block(o->deemph_fir[i], t2, ns, t2); /* :896 */
    else iirfilt_rrrf_execute_block(o->deemph[i], t2, ns, t2);                       /* :898 */
    if (o->cfg.lowpass) firfilt_rrrf_execute_block(o->audio_filt[i], t2, ns, t2);
    if (out->audio) memcpy(out->audio + (size_t)i * out->ld, t2, (size_t)ns * sizeof(float));
    if (out->pcm)
      for (unsigned k = 0; k < ns; k++) out->pcm[(size_t)i * out->ld + k] = to_s16_sat(t2[k]);
  }

  /* :910-913 -- note the waterfall sees the UN-mixed resampler output (local array, :911) */
  if (o->asgram) {
    float pv = 0, pf = 0;
    char *ascii = out->ascii ? out->ascii : (char *)alloca(o->cfg.waterfall + 1);
    asgramcf_write(o->asgram, o->resamp_tmp, ny);
    asgramcf_execute(o->asgram, ascii, &pv, &pf);
    if (out->peak) { out->peak[0] = pv; out->peak[1] = pf; }
    if (out->psd) {
      unsigned np;
      const float *p = oracle_asgramcf_last_psd(o->asgram, &np);
      memcpy(out->psd, p, np * sizeof(float));
    }
  }
  *ny_out = ny;
  *ns_out = ns;
  return 0;
}

/* ------------------------------------------------------------------ dsd_in chain */
struct oracle_dsd_s {
  oracle_dsd_cfg cfg;
  iirfilt_crcf dcblock;
  msresamp_crcf res_down;
  msresamp_rrrf res_up;
  freqdem fm_demod;
  cf *buffp, *resamp_buf;
  float *fm_out, *out_buf;
  unsigned res_size, out_size;
};

void oracle_dsd_default_cfg(oracle_dsd_cfg *c) {
  memset(c, 0, sizeof(*c));
  c->fs_in = 1024000;      /* include/dsd_in.h:11 */
  c->in_fmt = ORACLE_FMT_CF32;
  c->fs_sig = 12500;       /* src/dsd_in.c:23 */
  c->fs_audio = 48000;     /* :22 */
  c->chunk = 200000;       /* :25 */
  c->dc_alpha = 0.0005f;   /* :97 */
  c->resamp_as = 60.0f;
  c->kf = 0.5f;
}

oracle_dsd *oracle_dsd_create(const oracle_dsd_cfg *cfg) {
  oracle_dsd *o = (oracle_dsd *)calloc(1, sizeof(*o));
  o->cfg = *cfg;
  o->dcblock = iirfilt_crcf_create_dc_blocker(cfg->dc_alpha);
  o->res_down = msresamp_crcf_create(((float)cfg->fs_sig) / cfg->fs_in, cfg->resamp_as);
  o->res_up = msresamp_rrrf_create(((float)cfg->fs_audio) / cfg->fs_sig, cfg->resamp_as);
  o->fm_demod = freqdem_create(cfg->kf);
  o->res_size = (unsigned)ceilf(1 + 2 * cfg->chunk * ((float)cfg->fs_sig / cfg->fs_in)); /* :137 */
  o->out_size = (unsigned)ceilf(1 + 2 * o->res_size * ((float)cfg->fs_audio / cfg->fs_sig)); /* :138 */
  o->buffp = (cf *)malloc((size_t)cfg->chunk * sizeof(cf));
  o->resamp_buf = (cf *)malloc((size_t)o->res_size * sizeof(cf));
  o->fm_out = (float *)malloc((size_t)o->res_size * sizeof(float));
  o->out_buf = (float *)malloc((size_t)o->out_size * sizeof(float));
  return o;
}
void oracle_dsd_destroy(oracle_dsd *o) {
  if (!o) return;
  freqdem_destroy(o->fm_demod);
  msresamp_rrrf_destroy(o->res_up);
  msresamp_crcf_destroy(o->res_down);
  iirfilt_crcf_destroy(o->dcblock);
  free(o->buffp); free(o->resamp_buf); free(o->fm_out); free(o->out_buf);
  free(o);
}
unsigned oracle_dsd_res_size(const oracle_dsd *o) { return o->res_size; }
unsigned oracle_dsd_out_size(const oracle_dsd *o) { return o->out_size; }

int oracle_dsd_execute(oracle_dsd *o, const void *iq, unsigned n, cf *res, float *fm, float *audio, int16_t *pcm, unsigned *ny_out, unsigned *nz_out) {
  if (n > o->cfg.chunk) return -1;
  unsigned ny = 0, nz = 0;
  convert_in(o->cfg.in_fmt, iq, n, o->buffp);
  iirfilt_crcf_execute_block(o->dcblock, o->buffp, n, o->buffp);           /* :167 */
  msresamp_crcf_execute(o->res_down, o->buffp, n, o->resamp_buf, &ny);     /* :168 */
  freqdem_demodulate_block(o->fm_demod, o->resamp_buf, ny, o->fm_out);     /* :169 */
  msresamp_rrrf_execute(o->res_up, o->fm_out, ny, o->out_buf, &nz);        /* :170 */
  if (res) memcpy(res, o->resamp_buf, (size_t)ny * sizeof(cf));
  if (fm) memcpy(fm, o->fm_out, (size_t)ny * sizeof(float));
  if (audio) memcpy(audio, o->out_buf, (size_t)nz * sizeof(float));
  if (pcm)
    for (unsigned i = 0; i < nz; i++) pcm[i] = to_s16(o->out_buf[i]);     /* :172-175 */
  *ny_out = ny;
  *nz_out = nz;
  return 0;
}
