/*
 * host/pmr446_liquid_loop.c -- the reference's PMR446 processing loop over a capture FILE, written
 * against the liquid-dsp object API only (liquid-signature tier of the boundary).
 *
 * Structure follows init_liquid() and the body of main()'s while-loop in
 * /root/reference/src/sdr_pmr446.c (:420-480, :788-913): the SoapySDR read (:789) is replaced by
 * fread(), RtAudio (:903-906) by an s16 file writer, and the squelch selector (:828-874) by a
 * fixed `-c channel` choice (the selector is control logic outside the DSP path).
 * The same source links against either
 *     libpmr446_b200.so   (include/pmr446_liquid_shim.h: every *_execute* runs on the GPU), or
 *     liboracle_pmr446.so (oracle/liquid_subset.h: CPU restatement, test infrastructure)
 * which is how tests/test_gpu_host_harness.py checks the drop-in claim at source level.
 *
 * usage: pmr446_liquid_loop -i capture.cf32 -o audio.s16 [-c chan(1..16)] [-g audio_gain] [-l] [-w W]
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef USE_ORACLE
#include "../oracle/liquid_subset.h"
#else
#include "../include/pmr446_liquid_shim.h"
#endif
#include "../include/pmr446_taps.h"

#define SDR_SAMPLERATE (1024000UL)
#define CHANNEL_WIDTH_HZ (12500UL)
#define NUM_CHANNELS (16)
#define SDR_RESAMPLERATE (NUM_CHANNELS * CHANNEL_WIDTH_HZ)
#define SDR_INPUT_CHUNK (100000UL)
#define SDR_RESAMP_BUF_SIZE (39064)
#define SDR_CHANNEL_BUF_SIZE (2441UL)

typedef struct {
  iirfilt_crcf dcblock;
  msresamp_crcf resampler;
  nco_crcf nco;
  firpfbch_crcf channelizer;
  freqdem fm_demod;
  firfilt_rrrf ctcss_filt;
  wdelayf ctcss_lp_delay;
  firfilt_rrrf audio_filt;
  iirfilt_rrrf deemph;
  cbuffercf resamp_buf;
  asgramcf asgram;
} chain_t;

#define CHECK(x) do { if (!(x)) { fprintf(stderr, "%s:%d: check failed: %s\n", __FILE__, __LINE__, #x); exit(2); } } while (0)

static void init_chain(chain_t *c, unsigned waterfall) {
  static float hp[PMR446_HP_AUDIO_TAPS_LEN], lp[PMR446_LP_AUDIO_TAPS_LEN];
  pmr446_hp_audio_taps_fill(hp);
  pmr446_lp_audio_taps_fill(lp);
  CHECK(c->dcblock = iirfilt_crcf_create_dc_blocker(0.0005f));
  CHECK(c->resampler = msresamp_crcf_create(((float)SDR_RESAMPLERATE) / SDR_SAMPLERATE, 60.0f));
  msresamp_crcf_print(c->resampler);
  CHECK(c->nco = nco_crcf_create(LIQUID_VCO));
  float offset = -0.5f * (float)(NUM_CHANNELS - 1) / (float)NUM_CHANNELS * 2 * M_PI;
  nco_crcf_set_frequency(c->nco, offset);
  CHECK(c->channelizer = firpfbch_crcf_create_kaiser(LIQUID_ANALYZER, NUM_CHANNELS, 13, 80.0));
  CHECK(c->fm_demod = freqdem_create(0.5f));
  CHECK(c->ctcss_filt = firfilt_rrrf_create(hp, PMR446_HP_AUDIO_TAPS_LEN));
  CHECK(c->ctcss_lp_delay = wdelayf_create((PMR446_HP_AUDIO_TAPS_LEN - 1) / 2));
  CHECK(c->audio_filt = firfilt_rrrf_create(lp, PMR446_LP_AUDIO_TAPS_LEN));
  CHECK(c->deemph = iirfilt_rrrf_create((float[]){PMR446_DEEMPH_B0, PMR446_DEEMPH_B1}, 2, (float[]){PMR446_DEEMPH_A0, PMR446_DEEMPH_A1}, 2));
  CHECK(c->resamp_buf = cbuffercf_create(SDR_RESAMP_BUF_SIZE));
  c->asgram = NULL;
  if (waterfall > 0) {
    CHECK(c->asgram = asgramcf_create(waterfall));
    asgramcf_set_scale(c->asgram, -40.0f, 2.0f);
  }
}

int main(int argc, char **argv) {
  const char *in = NULL, *out = NULL;
  int chan = 2, lowpass = 0;
  unsigned waterfall = 0;
  float audio_gain = 1.0f;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-i") && i + 1 < argc) in = argv[++i];
    else if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
    else if (!strcmp(argv[i], "-c") && i + 1 < argc) chan = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-g") && i + 1 < argc) audio_gain = (float)atof(argv[++i]);
    else if (!strcmp(argv[i], "-w") && i + 1 < argc) waterfall = (unsigned)atoi(argv[++i]);
    else if (!strcmp(argv[i], "-l")) lowpass = 1;
    else { fprintf(stderr, "usage: %s -i capture.cf32 -o audio.s16 [-c chan] [-g gain] [-l] [-w W]\n", argv[0]); return 1; }
  }
  if (!in || !out || chan < 1 || chan > NUM_CHANNELS) { fprintf(stderr, "bad arguments\n"); return 1; }
  FILE *fi = fopen(in, "rb"), *fo = fopen(out, "wb");
  CHECK(fi && fo);

  chain_t chain;
  init_chain(&chain, waterfall);
  const int active_chan = chan - 1;

  static float complex buffp[SDR_INPUT_CHUNK];
  static float complex resamp_buf[SDR_RESAMP_BUF_SIZE];
  static float complex chan_bufs[NUM_CHANNELS][SDR_CHANNEL_BUF_SIZE];
  static float tmp_buf1[SDR_CHANNEL_BUF_SIZE], tmp_buf2[SDR_CHANNEL_BUF_SIZE];
  static int16_t pcm[SDR_CHANNEL_BUF_SIZE];
  float complex tmp_chan_buf_out[NUM_CHANNELS];
  char ascii[2049];
  unsigned long total = 0;

  for (;;) {
    size_t rd = fread(buffp, sizeof(float complex), SDR_INPUT_CHUNK, fi);   /* was SoapySDRDevice_readStream, :789 */
    if (rd == 0) break;
    unsigned int ny = 0;
    CHECK(iirfilt_crcf_execute_block(chain.dcblock, (liquid_float_complex *)buffp, (unsigned)rd, (liquid_float_complex *)buffp) == LIQUID_OK);
    CHECK(msresamp_crcf_execute(chain.resampler, (liquid_float_complex *)buffp, (unsigned)rd, (liquid_float_complex *)resamp_buf, &ny) == LIQUID_OK);
    CHECK(cbuffercf_write(chain.resamp_buf, (liquid_float_complex *)resamp_buf, ny) == LIQUID_OK);

    size_t ns = 0;
    unsigned int num_read;
    liquid_float_complex *rpc;
    while (cbuffercf_size(chain.resamp_buf) >= NUM_CHANNELS) {
      cbuffercf_read(chain.resamp_buf, NUM_CHANNELS, &rpc, &num_read);
      CHECK(num_read == NUM_CHANNELS);
#ifdef PER_SAMPLE_NCO
      for (int i = 0; i < NUM_CHANNELS; i++) {          /* literal form of :808-812 */
        nco_crcf_mix_down(chain.nco, rpc[i], &rpc[i]);
        nco_crcf_step(chain.nco);
      }
#else
      nco_crcf_mix_block_down(chain.nco, rpc, rpc, NUM_CHANNELS);   /* liquid's block form of the same loop */
#endif
      CHECK(firpfbch_crcf_analyzer_execute(chain.channelizer, rpc, (liquid_float_complex *)tmp_chan_buf_out) == LIQUID_OK);
      CHECK(cbuffercf_release(chain.resamp_buf, num_read) == LIQUID_OK);
      for (size_t i = 0; i < NUM_CHANNELS; i++) chan_bufs[i][ns] = tmp_chan_buf_out[i];
      ns++;
    }
    CHECK(ns <= SDR_CHANNEL_BUF_SIZE);

    for (int i = 0; i < NUM_CHANNELS; i++) {
      if (active_chan == i) {
        float tmp;
        freqdem_demodulate_block(chain.fm_demod, (liquid_float_complex *)chan_bufs[i], (unsigned)ns, tmp_buf1);
        firfilt_rrrf_execute_block(chain.ctcss_filt, tmp_buf1, (unsigned)ns, tmp_buf2);
        for (size_t k = 0; k < ns; k++) {
          wdelayf_push(chain.ctcss_lp_delay, tmp_buf1[k]);
          wdelayf_read(chain.ctcss_lp_delay, &tmp);
          tmp_buf1[k] = tmp - tmp_buf2[k];
          tmp_buf2[k] *= audio_gain;
        }
        iirfilt_rrrf_execute_block(chain.deemph, tmp_buf2, (unsigned)ns, tmp_buf2);
        if (lowpass) firfilt_rrrf_execute_block(chain.audio_filt, tmp_buf2, (unsigned)ns, tmp_buf2);
        for (size_t k = 0; k < ns; k++) {   /* src/dsd_in.c:172-175, saturated like the device clips RtAudio's float audio */
          float y = tmp_buf2[k] * (float)INT16_MAX;
          y = y < -32768.0f ? -32768.0f : (y > 32767.0f ? 32767.0f : y);
          pcm[k] = (int16_t)(int32_t)y;
        }
        CHECK(fwrite(pcm, 2, ns, fo) == ns);   /* was cbufferf_write to RtAudio, :903-906 */
        total += ns;
      }
    }
    if (chain.asgram) {
      float maxval, maxfreq;
      asgramcf_write(chain.asgram, (liquid_float_complex *)resamp_buf, ny);
      asgramcf_execute(chain.asgram, ascii, &maxval, &maxfreq);
      ascii[waterfall] = '\0';
      printf(" > %s < pk%5.1fdB [%5.2f]\n", ascii, maxval, maxfreq);
    }
  }
  fprintf(stderr, "wrote %lu audio samples of channel %d\n", total, chan);
  fclose(fi);
  fclose(fo);
  return 0;
}
