/*
 * oracle/receiver.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see liquid_subset.h).
 *
 * CPU restatement of what the reference does with the channelizer output of each chunk
 * (/root/reference/src/sdr_pmr446.c):
 *   average_power()            :330-336   20 log10(mean |x|) per channel
 *   find_max_rssi_channel()    :668-700   strongest enabled channel, "RSSI" = max - mean [dB]
 *   state machine              :828-874   scanning <-> tuned with 5 dB hysteresis, lock modes
 *   selected channel chain     :876-908   ONE freqdem / FIR / delay / de-emphasis whose state
 *                                         carries over channel changes; freqdem_reset on detune
 *   ctcss_execute()            :605-628   DC blocker + detector on the complementary branch
 *   ctcss_detector_*()         :338-407   38-tone Goertzel bank over blocks of 2441 samples
 * The front half (DC block .. channelizer) is the oracle_pmr object with channelize_only = 1.
 */
#include "receiver.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pmr446_taps.h"

typedef float _Complex cf;

typedef struct {
  float coef[ORACLE_CTCSS_NUM_FREQS];
  float u0[ORACLE_CTCSS_NUM_FREQS], u1[ORACLE_CTCSS_NUM_FREQS], power[ORACLE_CTCSS_NUM_FREQS];
  float max_power;
  int max_power_index;
  unsigned samp_processed;
  int tone_detected;
} tone_bank;

struct oracle_rx_s {
  oracle_rx_cfg cfg;
  oracle_pmr *front;
  unsigned M, chan_size;
  int state, active_chan;
  float rssi, ctcss_freq;
  freqdem fm_demod;
  firfilt_rrrf ctcss_filt, audio_filt;
  wdelayf ctcss_lp_delay;
  iirfilt_rrrf deemph, ctcss_dcblock;
  tone_bank tones;
  cf *chan;
  float *t1, *t2, *rssi_ch;
};

void oracle_rx_default_cfg(oracle_rx_cfg *c) {
  memset(c, 0, sizeof(*c));
  oracle_pmr_default_cfg(&c->chain);
  c->squelch_level = 18.0f;
  c->channel_mask = ~0ull;
  c->lock_mode = 0;
  c->ctcss_block = 2441;
  c->ctcss_dc_alpha = 0.0005f;
}

/* ctcss_detector_reset, :338-347 */
static void tones_reset(tone_bank *t) {
  t->samp_processed = 0;
  t->max_power = 0.0f;
  t->max_power_index = 0;
  t->tone_detected = 0;
  for (int j = 0; j < ORACLE_CTCSS_NUM_FREQS; j++) t->power[j] = t->u0[j] = t->u1[j] = 0.0f;
}

oracle_rx *oracle_rx_create(const oracle_rx_cfg *cfg) {
  if (cfg->chain.num_channels > 64) return NULL; /* MAX_CHANNELS, :18; the mask is 64 bits wide */
  oracle_rx *o = (oracle_rx *)calloc(1, sizeof(*o));
  o->cfg = *cfg;
  o->cfg.chain.channelize_only = 1;
  o->cfg.chain.active_only = -1;
  o->front = oracle_pmr_create(&o->cfg.chain);
  if (!o->front) { free(o); return NULL; }
  o->M = cfg->chain.num_channels;
  o->chan_size = oracle_pmr_chan_size(o->front);
  o->state = 0;         /* proc_scanning, :147 */
  o->active_chan = -1;  /* :148 */
  o->ctcss_freq = -1.0f; /* :149 */
  float hp[PMR446_HP_AUDIO_TAPS_LEN], lp[PMR446_LP_AUDIO_TAPS_LEN];
  pmr446_hp_audio_taps_fill(hp);
  pmr446_lp_audio_taps_fill(lp);
  o->fm_demod = freqdem_create(cfg->chain.kf);
  o->ctcss_filt = firfilt_rrrf_create(hp, PMR446_HP_AUDIO_TAPS_LEN);
  o->ctcss_lp_delay = wdelayf_create((PMR446_HP_AUDIO_TAPS_LEN - 1) / 2);
  o->ctcss_dcblock = iirfilt_rrrf_create_dc_blocker(cfg->ctcss_dc_alpha);
  o->audio_filt = firfilt_rrrf_create(lp, PMR446_LP_AUDIO_TAPS_LEN);
  o->deemph = iirfilt_rrrf_create((float[]){PMR446_DEEMPH_B0, PMR446_DEEMPH_B1}, 2, (float[]){PMR446_DEEMPH_A0, PMR446_DEEMPH_A1}, 2);
  tones_reset(&o->tones);
  const double fs_audio = (double)cfg->chain.channel_width; /* AUDIO_SAMPLERATE == CHANNEL_WIDTH_HZ, :24 */
  for (int j = 0; j < ORACLE_CTCSS_NUM_FREQS; j++) o->tones.coef[j] = 2.0f * cosf((2.0 * M_PI * pmr446_ctcss_freqs[j]) / fs_audio); /* :360-361 */
  o->chan = (cf *)malloc((size_t)o->M * o->chan_size * sizeof(cf));
  o->t1 = (float *)malloc((size_t)o->chan_size * sizeof(float));
  o->t2 = (float *)malloc((size_t)o->chan_size * sizeof(float));
  o->rssi_ch = (float *)malloc((size_t)o->M * sizeof(float));
  return o;
}

void oracle_rx_destroy(oracle_rx *o) {
  if (!o) return;
  oracle_pmr_destroy(o->front);
  freqdem_destroy(o->fm_demod);
  firfilt_rrrf_destroy(o->ctcss_filt);
  wdelayf_destroy(o->ctcss_lp_delay);
  iirfilt_rrrf_destroy(o->ctcss_dcblock);
  firfilt_rrrf_destroy(o->audio_filt);
  iirfilt_rrrf_destroy(o->deemph);
  free(o->chan); free(o->t1); free(o->t2); free(o->rssi_ch);
  free(o);
}

unsigned oracle_rx_chan_size(const oracle_rx *o) { return o->chan_size; }

/* :330-336 */
static float mean_abs_db(const cf *x, unsigned len) {
  float acc = 0.0f;
  for (unsigned i = 0; i < len; i++) acc += cabsf(x[i]);
  return 20 * log10f(acc / len);
}

/* :668-700; *spread = strongest - mean over the enabled channels */
static int strongest_channel(const oracle_rx *o, const float *rssi_ch, float *spread) {
  int best = -1, enabled = 0;
  float top = 0.0f, sum = 0.0f;
  for (unsigned i = 0; i < o->M; i++) {
    if (!(o->cfg.channel_mask & (1ull << i))) continue;
    enabled++;
    sum += rssi_ch[i];
    if (best < 0 || rssi_ch[i] > top) { top = rssi_ch[i]; best = (int)i; }
  }
  if (best >= 0) *spread = top - sum / enabled;
  return best;
}

/* ctcss_detector_analyze, :365-407 */
static void tones_analyze(tone_bank *t, const float *x, unsigned n, unsigned block) {
  for (unsigned i = 0; i < n; i++) {
    for (int j = 0; j < ORACLE_CTCSS_NUM_FREQS; j++) {
      const float older = t->u0[j];
      t->u0[j] = x[i] + (t->coef[j] * t->u0[j]) - t->u1[j];
      t->u1[j] = older;
    }
    if (++t->samp_processed != block) continue;
    float avg = 0.0f;
    t->max_power = 0.0f;
    for (int j = 0; j < ORACLE_CTCSS_NUM_FREQS; j++) {
      t->power[j] = (t->u0[j] * t->u0[j]) + (t->u1[j] * t->u1[j]) - (t->coef[j] * t->u0[j] * t->u1[j]);
      t->u0[j] = t->u1[j] = 0.0f;
      avg += t->power[j];
      if (t->power[j] > t->max_power) { t->max_power = t->power[j]; t->max_power_index = j; }
    }
    avg /= ORACLE_CTCSS_NUM_FREQS;
    t->tone_detected = (avg > 120.0f) && ((t->max_power / avg) > 10.0f);
    t->samp_processed = 0;
  }
}

int oracle_rx_execute(oracle_rx *o, const void *iq, unsigned n, const oracle_rx_out *out, oracle_rx_status *st, unsigned *ns_out) {
  oracle_pmr_out fo;
  memset(&fo, 0, sizeof fo);
  fo.chan = o->chan;
  fo.ld = o->chan_size;
  unsigned ny = 0, ns = 0;
  int rc = oracle_pmr_execute(o->front, iq, n, &fo, &ny, &ns);
  if (rc) return rc;
  if (out->ld && ns > out->ld) return -3;
  const unsigned M = o->M;
  for (unsigned i = 0; i < M; i++) o->rssi_ch[i] = mean_abs_db(o->chan + (size_t)i * o->chan_size, ns);
  if (out->rssi) memcpy(out->rssi, o->rssi_ch, M * sizeof(float));
  if (out->chan)
    for (unsigned i = 0; i < M; i++) memcpy(out->chan + (size_t)i * out->ld, o->chan + (size_t)i * o->chan_size, (size_t)ns * sizeof(cf));

  int events = 0;
  float spread = 0.0f;
  const int best = strongest_channel(o, o->rssi_ch, &spread);
  /* :828-874 */
  if (o->state == 0) {
    o->rssi = spread;
    if (o->rssi > o->cfg.squelch_level) {
      o->active_chan = best;
      o->state = 1;
      events |= 1;
    }
  } else {
    o->rssi = spread;
    if (o->cfg.lock_mode == 1 && o->active_chan != best) {
      o->active_chan = best;
      events |= 2;
    }
    if (o->rssi < (o->cfg.squelch_level - 5.0)) {
      o->active_chan = -1;
      o->state = 0;
      o->ctcss_freq = 0.0f;
      freqdem_reset(o->fm_demod);
      tones_reset(&o->tones);
      events |= 4;
    }
  }

  /* :876-908 for the selected channel */
  unsigned n_audio = 0;
  if (o->active_chan >= 0) {
    float *t1 = o->t1, *t2 = o->t2;
    freqdem_demodulate_block(o->fm_demod, o->chan + (size_t)o->active_chan * o->chan_size, ns, t1);
    firfilt_rrrf_execute_block(o->ctcss_filt, t1, ns, t2);
    for (unsigned k = 0; k < ns; k++) {
      float delayed;
      wdelayf_push(o->ctcss_lp_delay, t1[k]);
      wdelayf_read(o->ctcss_lp_delay, &delayed);
      t1[k] = delayed - t2[k];
      t2[k] *= o->cfg.chain.audio_gain;
    }
    /* ctcss_execute, :605-628 */
    iirfilt_rrrf_execute_block(o->ctcss_dcblock, t1, ns, t1);
    if (out->ctcss_in) memcpy(out->ctcss_in, t1, (size_t)ns * sizeof(float));
    const int had_tone = o->tones.tone_detected, had_code = o->tones.max_power_index;
    tones_analyze(&o->tones, t1, ns, o->cfg.ctcss_block);
    o->ctcss_freq = pmr446_ctcss_freqs[o->tones.max_power_index];
    if (o->tones.tone_detected) {
      if (!had_tone) events |= 8;
      else if (had_code != o->tones.max_power_index) events |= 16;
    } else if (had_tone) {
      events |= 32;
    }
    iirfilt_rrrf_execute_block(o->deemph, t2, ns, t2);
    if (o->cfg.chain.lowpass) firfilt_rrrf_execute_block(o->audio_filt, t2, ns, t2);
    if (out->audio) memcpy(out->audio, t2, (size_t)ns * sizeof(float));
    if (out->pcm)
      for (unsigned k = 0; k < ns; k++) {
        float y = t2[k] * (float)INT16_MAX;   /* saturating, see chains.c to_s16_sat */
        y = y < -32768.0f ? -32768.0f : (y > 32767.0f ? 32767.0f : y);
        out->pcm[k] = (int16_t)(int32_t)y;
      }
    n_audio = ns;
  }
  if (out->ctcss_power) memcpy(out->ctcss_power, o->tones.power, sizeof o->tones.power);
  if (st) {
    st->state = o->state;
    st->active_chan = o->active_chan;
    st->rssi = o->rssi;
    st->n_audio = n_audio;
    st->tone_detected = o->tones.tone_detected;
    st->ctcss_index = o->tones.max_power_index;
    st->ctcss_freq = o->ctcss_freq;
    st->max_power = o->tones.max_power;
    st->events = events;
  }
  if (ns_out) *ns_out = ns;
  return 0;
}
