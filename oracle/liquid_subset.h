/*
 * oracle/liquid_subset.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, float32, single thread per object) of the subset of the
 * liquid-dsp v1.7.0 object API that the reference application calls on its receive hot
 * path (/root/reference/src/sdr_pmr446.c:420-518, :788-931; /root/reference/src/dsd_in.c:95-124,
 * :159-180).  liquid-dsp itself is an un-vendored third-party dependency of the reference
 * (pinned at tag v1.7.0 by /root/reference/.github/workflows/build.yml:30, configured with
 * --enable-simdoverride, :33) and is absent from this build image, so the algorithms are
 * restated here from the library's published behaviour.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or fixtures
 * for this path, and neither the reference nor liquid-dsp can be compiled here.  The
 * restatement is pinned only by analytic known-answer tests (tests/test_oracle_kat.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or load this code.  The product (sdr_pmr446_b200/) never does.
 *
 * The function names and signatures are liquid's, so that loop bodies transcribed from the
 * reference's main() compile against this header unchanged.
 */
#ifndef ORACLE_LIQUID_SUBSET_H
#define ORACLE_LIQUID_SUBSET_H

#include <complex.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float _Complex liquid_float_complex;

#define LIQUID_OK 0
#define LIQUID_EICONFIG 3
#define LIQUID_EIRANGE 5
typedef int liquid_error_code;

typedef enum { LIQUID_NCO = 0, LIQUID_VCO = 1 } liquid_ncotype;
#define LIQUID_ANALYZER 0
#define LIQUID_SYNTHESIZER 1

/* knobs for the Appendix-A items flagged "verify first"; defaults follow SURVEY.md */
typedef struct {
  int   resamp_npfb;      /* arbitrary resampler filter-bank size, default 256 (older lineage: 64)   */
  int   resamp_fc_mode;   /* 0: fc = min(0.49, 0.515*rate_arb) (default); 1: fc = 0.4                */
  int   kaiser_r_mode;    /* 0: r = 2t/(n-1) (default); 1: r = 2t/n                                  */
  int   asgram_avg;       /* 0: column = max of the p=4 bins (default); 1: average of the 4 dB values */
  int   asgram_keep_buf;  /* 0: execute() resets the sample window too (default); 1: keeps it        */
} oracle_liquid_knobs;
oracle_liquid_knobs *oracle_liquid_get_knobs(void);

/* ---- filter design helpers (exposed for the known-answer tests) ---- */
float liquid_besseli0f(float z);
float kaiser_beta_As(float as);
float liquid_kaiser(unsigned i, unsigned wlen, float beta);
int   liquid_firdes_kaiser(unsigned n, float fc, float as, float mu, float *h);
unsigned estimate_req_filter_len(float df, float as);

/* ---- iirfilt (A.1) ---- */
typedef struct iirfilt_crcf_s *iirfilt_crcf;
typedef struct iirfilt_rrrf_s *iirfilt_rrrf;
iirfilt_crcf iirfilt_crcf_create_dc_blocker(float alpha);
int iirfilt_crcf_execute_block(iirfilt_crcf q, liquid_float_complex *x, unsigned n, liquid_float_complex *y);
int iirfilt_crcf_destroy(iirfilt_crcf q);
iirfilt_rrrf iirfilt_rrrf_create(float *b, unsigned nb, float *a, unsigned na);
iirfilt_rrrf iirfilt_rrrf_create_dc_blocker(float alpha);
int iirfilt_rrrf_execute_block(iirfilt_rrrf q, float *x, unsigned n, float *y);
int iirfilt_rrrf_destroy(iirfilt_rrrf q);

/* ---- msresamp (A.2-A.5) ---- */
typedef struct msresamp_s_crcf *msresamp_crcf;
typedef struct msresamp_s_rrrf *msresamp_rrrf;
msresamp_crcf msresamp_crcf_create(float rate, float as);
int msresamp_crcf_execute(msresamp_crcf q, liquid_float_complex *x, unsigned nx, liquid_float_complex *y, unsigned *ny);
int msresamp_crcf_print(msresamp_crcf q);
int msresamp_crcf_destroy(msresamp_crcf q);
msresamp_rrrf msresamp_rrrf_create(float rate, float as);
int msresamp_rrrf_execute(msresamp_rrrf q, float *x, unsigned nx, float *y, unsigned *ny);
int msresamp_rrrf_print(msresamp_rrrf q);
int msresamp_rrrf_destroy(msresamp_rrrf q);
/* introspection used by tests (not in liquid) */
int oracle_msresamp_crcf_plan(msresamp_crcf q, unsigned *stages, unsigned *m_stage /*[stages]*/, float *rate_arb, unsigned *step, unsigned *npfb);

/* ---- nco (A.7) ---- */
typedef struct nco_crcf_s *nco_crcf;
nco_crcf nco_crcf_create(liquid_ncotype type);
int nco_crcf_set_frequency(nco_crcf q, float dtheta);
int nco_crcf_mix_down(nco_crcf q, liquid_float_complex x, liquid_float_complex *y);
int nco_crcf_step(nco_crcf q);
int nco_crcf_mix_block_down(nco_crcf q, liquid_float_complex *x, liquid_float_complex *y, unsigned n);
int nco_crcf_destroy(nco_crcf q);
unsigned oracle_nco_crcf_get_dtheta_u32(nco_crcf q);

/* ---- firpfbch analyzer (A.8) ---- */
typedef struct firpfbch_crcf_s *firpfbch_crcf;
firpfbch_crcf firpfbch_crcf_create_kaiser(int type, unsigned M, unsigned m, float as);
int firpfbch_crcf_analyzer_execute(firpfbch_crcf q, liquid_float_complex *x, liquid_float_complex *y);
int firpfbch_crcf_destroy(firpfbch_crcf q);

/* ---- freqdem (A.9) ---- */
typedef struct freqdem_s *freqdem;
freqdem freqdem_create(float kf);
int freqdem_demodulate_block(freqdem q, liquid_float_complex *r, unsigned n, float *m);
int freqdem_reset(freqdem q);
int freqdem_destroy(freqdem q);

/* ---- firfilt (A.10) ---- */
typedef struct firfilt_rrrf_s *firfilt_rrrf;
firfilt_rrrf firfilt_rrrf_create(float *h, unsigned n);
int firfilt_rrrf_execute_block(firfilt_rrrf q, float *x, unsigned n, float *y);
int firfilt_rrrf_destroy(firfilt_rrrf q);

/* ---- wdelay (A.11) ---- */
typedef struct wdelayf_s *wdelayf;
wdelayf wdelayf_create(unsigned delay);
int wdelayf_push(wdelayf q, float v);
int wdelayf_read(wdelayf q, float *v);
int wdelayf_destroy(wdelayf q);

/* ---- cbuffer (A.12) ---- */
typedef struct cbuffercf_s *cbuffercf;
typedef struct cbufferf_s *cbufferf;
cbuffercf cbuffercf_create(unsigned max_size);
int cbuffercf_write(cbuffercf q, liquid_float_complex *v, unsigned n);
unsigned cbuffercf_size(cbuffercf q);
int cbuffercf_read(cbuffercf q, unsigned n, liquid_float_complex **v, unsigned *nr);
int cbuffercf_release(cbuffercf q, unsigned n);
int cbuffercf_destroy(cbuffercf q);
cbufferf cbufferf_create(unsigned max_size);
int cbufferf_write(cbufferf q, float *v, unsigned n);
unsigned cbufferf_size(cbufferf q);
unsigned cbufferf_max_size(cbufferf q);
int cbufferf_read(cbufferf q, unsigned n, float **v, unsigned *nr);
int cbufferf_release(cbufferf q, unsigned n);
int cbufferf_destroy(cbufferf q);

/* ---- asgram / spgram (A.13) ---- */
typedef struct asgramcf_s *asgramcf;
asgramcf asgramcf_create(unsigned nfft);
int asgramcf_set_scale(asgramcf q, float ref, float div);
int asgramcf_write(asgramcf q, liquid_float_complex *x, unsigned n);
int asgramcf_execute(asgramcf q, char *ascii, float *peakval, float *peakfreq);
int asgramcf_destroy(asgramcf q);
/* test hook: the nfft*4 dB values of the last execute() (fft-shifted), not in liquid */
const float *oracle_asgramcf_last_psd(asgramcf q, unsigned *n);

/* generic float32 mixed-radix FFT used by firpfbch and spgram (forward, unnormalised) */
void oracle_fft_forward(unsigned n, const liquid_float_complex *in, liquid_float_complex *out);

#ifdef __cplusplus
}
#endif
#endif
