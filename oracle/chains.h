/*
 * oracle/chains.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see liquid_subset.h).
 * Driver API of the two reference chains restated in oracle/chains.c.
 */
#ifndef ORACLE_CHAINS_H
#define ORACLE_CHAINS_H

#include <stdint.h>

#include "liquid_subset.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_FMT_CF32 0
#define ORACLE_FMT_CU8 1

typedef struct {
  unsigned fs_in;          /* input sample rate [Hz]                                   */
  int in_fmt;              /* ORACLE_FMT_*                                              */
  unsigned num_channels;   /* M                                                         */
  unsigned channel_width;  /* Hz; resampled rate = M * channel_width                    */
  unsigned pfb_m;          /* channelizer prototype semi-length                         */
  float pfb_as;            /* channelizer stop-band attenuation [dB]                    */
  float resamp_as;         /* msresamp stop-band attenuation [dB]                       */
  float dc_alpha;          /* DC blocker alpha                                          */
  float kf;                /* freqdem modulation factor                                 */
  float audio_gain;
  int lowpass;             /* apply the 103-tap audio low-pass                          */
  unsigned waterfall;      /* asgram width W (0 = off)                                  */
  unsigned chunk;          /* max samples per execute()                                 */
  int active_only;         /* -1: demodulate all channels; k: only channel k            */
  int channelize_only;     /* stop after the channelizer (timing of the front half)     */
  int deemph_fir;          /* APP_FIR_DEEMPH build: 101-tap FIR de-emphasis (:122-135, :458) */
} oracle_pmr_cfg;

/* per-chunk outputs; any pointer may be NULL.  Channel-major arrays use row stride ld. */
typedef struct {
  liquid_float_complex *dcblocked; /* [n]                                                */
  liquid_float_complex *res;       /* [ny]   un-mixed resampler output                   */
  liquid_float_complex *chan;      /* [M][ld] channelizer output                         */
  float *demod;                    /* [M][ld] discriminator output                       */
  float *lpcomp;                   /* [M][ld] delayed - high-passed (CTCSS branch input) */
  float *audio;                    /* [M][ld] final float audio                          */
  int16_t *pcm;                    /* [M][ld] (int16_t)(audio * 32767)                   */
  unsigned ld;
  char *ascii;                     /* [W] waterfall row                                  */
  float *peak;                     /* [2] peak value [dB], peak frequency [-0.5,0.5)     */
  float *psd;                      /* [4W] dB values behind the row                      */
} oracle_pmr_out;

typedef struct oracle_pmr_s oracle_pmr;
void oracle_pmr_default_cfg(oracle_pmr_cfg *c);
oracle_pmr *oracle_pmr_create(const oracle_pmr_cfg *cfg);
void oracle_pmr_destroy(oracle_pmr *o);
unsigned oracle_pmr_res_size(const oracle_pmr *o);
unsigned oracle_pmr_chan_size(const oracle_pmr *o);
int oracle_pmr_execute(oracle_pmr *o, const void *iq, unsigned n, const oracle_pmr_out *out, unsigned *ny, unsigned *ns);

typedef struct {
  unsigned fs_in;
  int in_fmt;
  unsigned fs_sig;    /* 12500 */
  unsigned fs_audio;  /* 48000 */
  unsigned chunk;
  float dc_alpha, resamp_as, kf;
} oracle_dsd_cfg;
typedef struct oracle_dsd_s oracle_dsd;
void oracle_dsd_default_cfg(oracle_dsd_cfg *c);
oracle_dsd *oracle_dsd_create(const oracle_dsd_cfg *cfg);
void oracle_dsd_destroy(oracle_dsd *o);
unsigned oracle_dsd_res_size(const oracle_dsd *o);
unsigned oracle_dsd_out_size(const oracle_dsd *o);
int oracle_dsd_execute(oracle_dsd *o, const void *iq, unsigned n, liquid_float_complex *res, float *fm, float *audio, int16_t *pcm,
                       unsigned *ny, unsigned *nz);

#ifdef __cplusplus
}
#endif
#endif
