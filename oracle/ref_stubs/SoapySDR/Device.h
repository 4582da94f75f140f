/* TEST INFRASTRUCTURE: stand-in for <SoapySDR/Device.h> declaring exactly what /root/reference/src/shared.c:11-87,
 * src/sdr_pmr446.c:789 and src/dsd_in.c:161 use; implemented over a capture file in ../stubs.c (see ../README.md). */
#pragma once
#include <stdbool.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
#define SOAPY_SDR_RX 1
typedef struct SoapySDRDevice SoapySDRDevice;
typedef struct SoapySDRStream SoapySDRStream;
typedef struct { size_t size; char **keys; char **vals; } SoapySDRKwargs;
typedef struct { double minimum, maximum, step; } SoapySDRRange;
SoapySDRKwargs *SoapySDRDevice_enumerate(const SoapySDRKwargs *args, size_t *length);
int SoapySDRKwargs_set(SoapySDRKwargs *args, const char *key, const char *val);
void SoapySDRKwargs_clear(SoapySDRKwargs *args);
void SoapySDRKwargsList_clear(SoapySDRKwargs *args, size_t length);
SoapySDRDevice *SoapySDRDevice_make(const SoapySDRKwargs *args);
int SoapySDRDevice_unmake(SoapySDRDevice *device);
SoapySDRRange *SoapySDRDevice_getFrequencyRange(const SoapySDRDevice *device, int direction, size_t channel, size_t *length);
size_t SoapySDRDevice_getNumChannels(const SoapySDRDevice *device, int direction);
int SoapySDRDevice_setSampleRate(SoapySDRDevice *device, int direction, size_t channel, double rate);
int SoapySDRDevice_setFrequency(SoapySDRDevice *device, int direction, size_t channel, double frequency, const SoapySDRKwargs *args);
int SoapySDRDevice_setGain(SoapySDRDevice *device, int direction, size_t channel, double value);
SoapySDRStream *SoapySDRDevice_setupStream(SoapySDRDevice *device, int direction, const char *format, const size_t *channels,
                                           size_t numChans, const SoapySDRKwargs *args);
int SoapySDRDevice_activateStream(SoapySDRDevice *device, SoapySDRStream *stream, int flags, long long timeNs, size_t numElems);
int SoapySDRDevice_deactivateStream(SoapySDRDevice *device, SoapySDRStream *stream, int flags, long long timeNs);
int SoapySDRDevice_closeStream(SoapySDRDevice *device, SoapySDRStream *stream);
int SoapySDRDevice_readStream(SoapySDRDevice *device, SoapySDRStream *stream, void *const *buffs, size_t numElems, int *flags,
                              long long *timeNs, long timeoutUs);
#ifdef __cplusplus
}
#endif
