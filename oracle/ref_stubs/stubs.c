/* TEST INFRASTRUCTURE, NOT PRODUCT CODE: file-backed implementations of the SoapySDR / RtAudio / dlg symbols the
 * reference uses, so that its UNMODIFIED sources (/root/reference/src/sdr_pmr446.c, dsd_in.c, shared.c) run on a
 * recorded capture (oracle/ref.mk, oracle/ref_stubs/README.md).
 *
 *   REF_IQ            capture file (required)            REF_IQ_FMT   "cf32" (default) or "cu8"
 *   REF_AUDIO         float32 file the "audio device" writes what the RtAudio callback delivers
 *   REF_AUDIO_COUNTS  text file: samples delivered per readStream call (= per processed chunk)
 */
#include <complex.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <SoapySDR/Device.h>
#include <liquid/liquid.h>
#include <rtaudio/rtaudio_c.h>

void logging_init(void) {}   /* replaces src/logging.c (dlg internals); LOG() goes through dlg/dlg.h's macros */

/* ---- RtAudio: a file writer driven from the SDR read ------------------------------------------------------------ */
struct rtaudio { rtaudio_cb_t cb; void *userdata; int open, running; FILE *out, *counts; };
static struct rtaudio g_dac;
static const rtaudio_api_t g_apis[1] = {RTAUDIO_API_DUMMY};

unsigned int rtaudio_get_num_compiled_apis(void) { return 1; }
const rtaudio_api_t *rtaudio_compiled_api(void) { return g_apis; }
const char *rtaudio_api_name(rtaudio_api_t api) { return api == RTAUDIO_API_DUMMY ? "dummy" : "unspecified"; }
rtaudio_t rtaudio_create(rtaudio_api_t api) { (void)api; memset(&g_dac, 0, sizeof g_dac); return &g_dac; }
void rtaudio_destroy(rtaudio_t a) { (void)a; }
rtaudio_api_t rtaudio_current_api(rtaudio_t a) { (void)a; return RTAUDIO_API_DUMMY; }
int rtaudio_device_count(rtaudio_t a) { (void)a; return 1; }
unsigned int rtaudio_get_device_id(rtaudio_t a, int i) { (void)a; return 100u + (unsigned)i; }
unsigned int rtaudio_get_default_output_device(rtaudio_t a) { (void)a; return 100u; }
rtaudio_device_info_t rtaudio_get_device_info(rtaudio_t a, unsigned int id) {
  rtaudio_device_info_t d;
  (void)a;
  memset(&d, 0, sizeof d);
  d.id = id;
  d.output_channels = 1;
  d.is_default_output = 1;
  snprintf(d.name, sizeof d.name, "file writer (%s)", getenv("REF_AUDIO") ? getenv("REF_AUDIO") : "discarded");
  return d;
}
void rtaudio_show_warnings(rtaudio_t a, int show) { (void)a; (void)show; }
rtaudio_error_t rtaudio_open_stream(rtaudio_t a, rtaudio_stream_parameters_t *op, rtaudio_stream_parameters_t *ip, rtaudio_format_t fmt,
                                    unsigned int rate, unsigned int *frames, rtaudio_cb_t cb, void *userdata, rtaudio_stream_options_t *opt,
                                    rtaudio_error_cb_t errcb) {
  (void)op; (void)ip; (void)rate; (void)frames; (void)opt; (void)errcb;
  if (fmt != RTAUDIO_FORMAT_FLOAT32) return RTAUDIO_ERROR_WARNING;
  a->cb = cb;
  a->userdata = userdata;
  a->open = 1;
  if (getenv("REF_AUDIO")) a->out = fopen(getenv("REF_AUDIO"), "wb");
  if (getenv("REF_AUDIO_COUNTS")) a->counts = fopen(getenv("REF_AUDIO_COUNTS"), "w");
  return RTAUDIO_ERROR_NONE;
}
rtaudio_error_t rtaudio_start_stream(rtaudio_t a) { a->running = 1; return RTAUDIO_ERROR_NONE; }
/* everything the main loop queued since the last call, through the reference's own callback (:520-544); userdata is the
 * cbufferf the reference handed to rtaudio_open_stream (:585), so cbufferf_size() says how many frames are real */
static void drain_audio(void) {
  static float buf[1 << 16];
  struct rtaudio *a = &g_dac;
  if (!a->running || !a->cb) return;
  unsigned int n = cbufferf_size((cbufferf)a->userdata);
  if (n > (1u << 16)) n = 1u << 16;
  if (n) a->cb(buf, NULL, n, 0.0, 0, a->userdata);
  if (a->out && n) fwrite(buf, sizeof(float), n, a->out);
  if (a->counts) fprintf(a->counts, "%u\n", n);
}
rtaudio_error_t rtaudio_stop_stream(rtaudio_t a) {
  drain_audio();
  a->running = 0;
  return RTAUDIO_ERROR_NONE;
}
int rtaudio_is_stream_open(rtaudio_t a) { return a->open; }
void rtaudio_close_stream(rtaudio_t a) {
  a->open = 0;
  if (a->out) fclose(a->out);
  if (a->counts) fclose(a->counts);
  a->out = a->counts = NULL;
}

/* ---- SoapySDR: one fake device reading a capture file ------------------------------------------------------------------ */
struct SoapySDRDevice { FILE *f; int cu8; unsigned long long calls; };
struct SoapySDRStream { int dummy; };
static struct SoapySDRDevice g_dev;
static struct SoapySDRStream g_stream;

static char *dupstr(const char *s) { char *d = malloc(strlen(s) + 1); strcpy(d, s); return d; }
int SoapySDRKwargs_set(SoapySDRKwargs *args, const char *key, const char *val) {
  args->keys = realloc(args->keys, (args->size + 1) * sizeof(char *));
  args->vals = realloc(args->vals, (args->size + 1) * sizeof(char *));
  args->keys[args->size] = dupstr(key);
  args->vals[args->size] = dupstr(val);
  args->size++;
  return 0;
}
void SoapySDRKwargs_clear(SoapySDRKwargs *args) {
  for (size_t i = 0; i < args->size; i++) { free(args->keys[i]); free(args->vals[i]); }
  free(args->keys);
  free(args->vals);
  memset(args, 0, sizeof *args);
}
void SoapySDRKwargsList_clear(SoapySDRKwargs *args, size_t length) {
  for (size_t i = 0; i < length; i++) SoapySDRKwargs_clear(&args[i]);
  free(args);
}
SoapySDRKwargs *SoapySDRDevice_enumerate(const SoapySDRKwargs *args, size_t *length) {
  (void)args;
  SoapySDRKwargs *r = calloc(1, sizeof *r);
  SoapySDRKwargs_set(r, "driver", "file");
  SoapySDRKwargs_set(r, "label", getenv("REF_IQ") ? getenv("REF_IQ") : "(REF_IQ not set)");
  *length = 1;
  return r;
}
SoapySDRDevice *SoapySDRDevice_make(const SoapySDRKwargs *args) {
  (void)args;
  const char *path = getenv("REF_IQ"), *fmt = getenv("REF_IQ_FMT");
  if (!path) { fprintf(stderr, "ref_stubs: REF_IQ is not set\n"); return NULL; }
  g_dev.f = fopen(path, "rb");
  if (!g_dev.f) { perror(path); return NULL; }
  g_dev.cu8 = fmt && strcmp(fmt, "cu8") == 0;
  return &g_dev;
}
int SoapySDRDevice_unmake(SoapySDRDevice *d) { if (d && d->f) { fclose(d->f); d->f = NULL; } return 0; }
SoapySDRRange *SoapySDRDevice_getFrequencyRange(const SoapySDRDevice *d, int dir, size_t ch, size_t *length) {
  (void)d; (void)dir; (void)ch;
  SoapySDRRange *r = calloc(1, sizeof *r);
  r->minimum = 24e6;
  r->maximum = 1766e6;
  *length = 1;
  return r;
}
size_t SoapySDRDevice_getNumChannels(const SoapySDRDevice *d, int dir) { (void)d; (void)dir; return 1; }
int SoapySDRDevice_setSampleRate(SoapySDRDevice *d, int dir, size_t ch, double rate) { (void)d; (void)dir; (void)ch; (void)rate; return 0; }
int SoapySDRDevice_setFrequency(SoapySDRDevice *d, int dir, size_t ch, double f, const SoapySDRKwargs *a) { (void)d; (void)dir; (void)ch; (void)f; (void)a; return 0; }
int SoapySDRDevice_setGain(SoapySDRDevice *d, int dir, size_t ch, double v) { (void)d; (void)dir; (void)ch; (void)v; return 0; }
SoapySDRStream *SoapySDRDevice_setupStream(SoapySDRDevice *d, int dir, const char *format, const size_t *chs, size_t n, const SoapySDRKwargs *a) {
  (void)d; (void)dir; (void)chs; (void)n; (void)a;
  if (strcmp(format, "CF32") != 0) return NULL;
  return &g_stream;
}
int SoapySDRDevice_activateStream(SoapySDRDevice *d, SoapySDRStream *s, int flags, long long t, size_t n) { (void)d; (void)s; (void)flags; (void)t; (void)n; return 0; }
int SoapySDRDevice_deactivateStream(SoapySDRDevice *d, SoapySDRStream *s, int flags, long long t) { (void)d; (void)s; (void)flags; (void)t; return 0; }
int SoapySDRDevice_closeStream(SoapySDRDevice *d, SoapySDRStream *s) { (void)d; (void)s; return 0; }

int SoapySDRDevice_readStream(SoapySDRDevice *d, SoapySDRStream *s, void *const *buffs, size_t numElems, int *flags, long long *timeNs,
                              long timeoutUs) {
  (void)s; (void)timeoutUs;
  if (flags) *flags = 0;
  if (timeNs) *timeNs = 0;
  drain_audio();   /* the audio of the chunk processed since the previous read */
  float complex *out = (float complex *)buffs[0];
  size_t got;
  if (d->cu8) {
    static uint8_t *raw;
    static size_t raw_cap;
    if (raw_cap < 2 * numElems) { raw = realloc(raw, 2 * numElems); raw_cap = 2 * numElems; }
    got = fread(raw, 2, numElems, d->f);
    for (size_t i = 0; i < got; i++)   /* SoapyRTLSDR's conversion */
      out[i] = ((float)raw[2 * i] - 127.4f) * (1.0f / 128.0f) + (((float)raw[2 * i + 1] - 127.4f) * (1.0f / 128.0f)) * I;
  } else {
    got = fread(out, sizeof(float complex), numElems, d->f);
  }
  d->calls++;
  if (got == 0) {
#ifdef APP_DSD_IN
    fflush(stdout);          /* dsd_in's loop is while (true), src/dsd_in.c:159: the capture's end is the program's end */
    exit(EXIT_SUCCESS);
#else
    raise(SIGTERM);          /* sdr_pmr446's handler sets exit_via_sig (:190-199); the error return skips the loop body */
    return -1;
#endif
  }
  return (int)got;
}
