/*
 * oracle/receiver.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see liquid_subset.h).
 * The reference's receiver logic around the DSP chain, restated on the CPU: per-chunk RSSI,
 * squelch / channel-selector state machine, the single squelch-selected demodulation chain and
 * the CTCSS tone detector (/root/reference/src/sdr_pmr446.c:330-418, :605-628, :668-700, :828-908).
 */
#ifndef ORACLE_RECEIVER_H
#define ORACLE_RECEIVER_H

#include "chains.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_CTCSS_NUM_FREQS 38

typedef struct {
  oracle_pmr_cfg chain;
  float squelch_level;           /* :34, :152 (18 dB) */
  unsigned long long channel_mask; /* :155 */
  int lock_mode;                 /* 0 = lock_mode_start, 1 = lock_mode_max (include/sdr_pmr446.h:23-27) */
  unsigned ctcss_block;          /* CTCSS_BLOCK_SIZE :46 (2441) */
  float ctcss_dc_alpha;          /* :450 (0.0005) */
} oracle_rx_cfg;

typedef struct {
  int state;            /* proc_scanning 0 / proc_tuned 1, after this chunk's update */
  int active_chan;      /* -1 or 0-based channel */
  float rssi;           /* chain->rssi: max - mean of the enabled channels' average_power, dB */
  unsigned n_audio;     /* audio samples appended by this chunk (0 or ns) */
  int tone_detected;
  int ctcss_index;      /* max_power_index */
  float ctcss_freq;     /* chain->ctcss_freq (-1 at start, 0 after a detune) */
  float max_power;
  int events;           /* bit 0 tuned, 1 channel changed, 2 detuned, 3 CTCSS acquired, 4 code change, 5 lost */
} oracle_rx_status;

typedef struct {
  float *rssi;          /* [M] average_power per channel */
  float *audio;         /* [ld] */
  int16_t *pcm;         /* [ld] */
  float *ctcss_in;      /* [ld] CTCSS branch after its DC blocker */
  float *ctcss_power;   /* [38] Goertzel powers of the last finished block */
  liquid_float_complex *chan; /* [M][ld] */
  unsigned ld;
} oracle_rx_out;

typedef struct oracle_rx_s oracle_rx;
void oracle_rx_default_cfg(oracle_rx_cfg *c);
oracle_rx *oracle_rx_create(const oracle_rx_cfg *cfg);
void oracle_rx_destroy(oracle_rx *o);
unsigned oracle_rx_chan_size(const oracle_rx *o);
int oracle_rx_execute(oracle_rx *o, const void *iq, unsigned n, const oracle_rx_out *out, oracle_rx_status *st, unsigned *ns);

#ifdef __cplusplus
}
#endif
#endif
