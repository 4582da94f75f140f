/*
 * pmr446_liquid_shim.h -- liquid-dsp-signature tier of the drop-in boundary.
 *
 * Declares, with liquid-dsp v1.7.0's names and signatures, exactly the objects and functions the
 * reference calls on its receive path, so that the bodies of init_liquid()/destroy_liquid()/main()
 * in /root/reference/src/sdr_pmr446.c (:420-518, :795-913) and /root/reference/src/dsd_in.c
 * (:95-124, :167-170) compile against libpmr446_b200.so instead of libliquid.  Each declaration
 * cites the reference call site it serves.
 *
 * Every *_execute* / *_block / mix / analyzer call runs its arithmetic in a CUDA kernel on the
 * current device (batch of one stream): the caller's host buffer is copied to the GPU, processed,
 * and copied back before the call returns.  Filter state lives on the device between calls.
 * There is no CPU fallback: without a CUDA device every *_create() returns NULL.
 * cbuffer{cf,f} and wdelayf hold no arithmetic (a FIFO and a delay line); they are plain host
 * containers, as is the handle bookkeeping.
 *
 * Per-sample entry points (nco_crcf_mix_down, firpfbch_crcf_analyzer_execute, wdelayf_push) are
 * provided for literal source compatibility but cost one kernel launch per call; a port of the
 * reference should use the coarse calls of pmr446_b200.h, which replace the whole loop bodies
 * (see INTEGRATION.md).  All functions returning int return 0 (LIQUID_OK) on success.
 */
#ifndef PMR446_LIQUID_SHIM_H
#define PMR446_LIQUID_SHIM_H

#include <complex.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef LIQUID_OK
#define LIQUID_OK 0
#endif
#ifndef LIQUID_EICONFIG
#define LIQUID_EICONFIG 3
#endif
#ifndef LIQUID_EIRANGE
#define LIQUID_EIRANGE 5
#endif
typedef int liquid_error_code;
#ifdef __cplusplus
typedef struct { float re, im; } liquid_float_complex;   /* layout of C99 float _Complex */
#else
typedef float _Complex liquid_float_complex;
#endif
typedef enum { LIQUID_NCO = 0, LIQUID_VCO = 1 } liquid_ncotype;
#define LIQUID_ANALYZER 0
#define LIQUID_SYNTHESIZER 1

/* iirfilt_crcf: src/sdr_pmr446.c:422,795,516; src/dsd_in.c:97,167,122 */
typedef struct iirfilt_crcf_s *iirfilt_crcf;
iirfilt_crcf iirfilt_crcf_create_dc_blocker(float alpha);
int iirfilt_crcf_execute_block(iirfilt_crcf q, liquid_float_complex *x, unsigned int n, liquid_float_complex *y);
int iirfilt_crcf_destroy(iirfilt_crcf q);

/* iirfilt_rrrf: src/sdr_pmr446.c:450,461-463,606,898,497,502 (orders up to 1, which is all the reference uses) */
typedef struct iirfilt_rrrf_s *iirfilt_rrrf;
iirfilt_rrrf iirfilt_rrrf_create(float *b, unsigned int nb, float *a, unsigned int na);
iirfilt_rrrf iirfilt_rrrf_create_dc_blocker(float alpha);
int iirfilt_rrrf_execute_block(iirfilt_rrrf q, float *x, unsigned int n, float *y);
int iirfilt_rrrf_destroy(iirfilt_rrrf q);

/* msresamp_crcf: src/sdr_pmr446.c:425-428,796,514; src/dsd_in.c:100,168,120 (decimating rates) */
typedef struct msresamp_crcf_s *msresamp_crcf;
msresamp_crcf msresamp_crcf_create(float rate, float as);
int msresamp_crcf_execute(msresamp_crcf q, liquid_float_complex *x, unsigned int nx, liquid_float_complex *y, unsigned int *ny);
int msresamp_crcf_print(msresamp_crcf q);
int msresamp_crcf_destroy(msresamp_crcf q);

/* msresamp_rrrf: src/dsd_in.c:104,170 (interpolating rates in (2, 4]) */
typedef struct msresamp_rrrf_s *msresamp_rrrf;
msresamp_rrrf msresamp_rrrf_create(float rate, float as);
int msresamp_rrrf_execute(msresamp_rrrf q, float *x, unsigned int nx, float *y, unsigned int *ny);
int msresamp_rrrf_print(msresamp_rrrf q);
int msresamp_rrrf_destroy(msresamp_rrrf q);

/* nco_crcf: src/sdr_pmr446.c:430-434,808-812,512 */
typedef struct nco_crcf_s *nco_crcf;
nco_crcf nco_crcf_create(liquid_ncotype type);
int nco_crcf_set_frequency(nco_crcf q, float dtheta);
int nco_crcf_mix_down(nco_crcf q, liquid_float_complex x, liquid_float_complex *y);
int nco_crcf_step(nco_crcf q);
/* liquid's block form: y[i] = x[i] * conj(exp(j theta_i)), stepping once per sample */
int nco_crcf_mix_block_down(nco_crcf q, liquid_float_complex *x, liquid_float_complex *y, unsigned int n);
int nco_crcf_destroy(nco_crcf q);

/* firpfbch_crcf: src/sdr_pmr446.c:436-437,814,510 */
typedef struct firpfbch_crcf_s *firpfbch_crcf;
firpfbch_crcf firpfbch_crcf_create_kaiser(int type, unsigned int M, unsigned int m, float as);
int firpfbch_crcf_analyzer_execute(firpfbch_crcf q, liquid_float_complex *x, liquid_float_complex *y);
int firpfbch_crcf_destroy(firpfbch_crcf q);

/* freqdem: src/sdr_pmr446.c:440,881,866,508; src/dsd_in.c:108,169,118 */
typedef struct freqdem_s *freqdem;
freqdem freqdem_create(float kf);
int freqdem_demodulate_block(freqdem q, liquid_float_complex *r, unsigned int n, float *m);
int freqdem_reset(freqdem q);
int freqdem_destroy(freqdem q);

/* firfilt_rrrf: src/sdr_pmr446.c:443-444,453-454,882,901,500,506 (in-place x == y allowed) */
typedef struct firfilt_rrrf_s *firfilt_rrrf;
firfilt_rrrf firfilt_rrrf_create(float *h, unsigned int n);
int firfilt_rrrf_execute_block(firfilt_rrrf q, float *x, unsigned int n, float *y);
int firfilt_rrrf_destroy(firfilt_rrrf q);

/* wdelayf: src/sdr_pmr446.c:447,885-887,504 (host container) */
typedef struct wdelayf_s *wdelayf;
wdelayf wdelayf_create(unsigned int delay);
int wdelayf_push(wdelayf q, float v);
int wdelayf_read(wdelayf q, float *v);
int wdelayf_destroy(wdelayf q);

/* cbuffercf / cbufferf: src/sdr_pmr446.c:467,470,797,804-805,815,530,539,904,923,927,490-492 (host containers) */
typedef struct cbuffercf_s *cbuffercf;
typedef struct cbufferf_s *cbufferf;
cbuffercf cbuffercf_create(unsigned int max_size);
int cbuffercf_write(cbuffercf q, liquid_float_complex *v, unsigned int n);
unsigned int cbuffercf_size(cbuffercf q);
int cbuffercf_read(cbuffercf q, unsigned int n, liquid_float_complex **v, unsigned int *nr);
int cbuffercf_release(cbuffercf q, unsigned int n);
int cbuffercf_destroy(cbuffercf q);
cbufferf cbufferf_create(unsigned int max_size);
int cbufferf_write(cbufferf q, float *v, unsigned int n);
unsigned int cbufferf_size(cbufferf q);
unsigned int cbufferf_max_size(cbufferf q);
int cbufferf_read(cbufferf q, unsigned int n, float **v, unsigned int *nr);
int cbufferf_release(cbufferf q, unsigned int n);
int cbufferf_destroy(cbufferf q);

/* asgramcf: src/sdr_pmr446.c:474-476,911-912,486.
 * Limits of this implementation (liquid has none): nfft (the terminal width) in [2, 2048]; at most 2^18 samples may be
 * written between two asgramcf_execute() calls (asgramcf_write returns LIQUID_EIRANGE beyond that; the reference writes
 * <= 39 064 per execute, :911-912).  asgramcf_create() returns NULL outside the width range; pmr446_last_error() says why. */
typedef struct asgramcf_s *asgramcf;
asgramcf asgramcf_create(unsigned int nfft);
int asgramcf_set_scale(asgramcf q, float ref, float div);
int asgramcf_write(asgramcf q, liquid_float_complex *x, unsigned int n);
int asgramcf_execute(asgramcf q, char *ascii, float *peakval, float *peakfreq);
int asgramcf_destroy(asgramcf q);

#ifdef __cplusplus
}
#endif
#endif
