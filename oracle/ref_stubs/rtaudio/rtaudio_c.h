/* TEST INFRASTRUCTURE: stand-in for <rtaudio/rtaudio_c.h> (RtAudio 6.0.0 C API) declaring exactly what
 * /root/reference/src/sdr_pmr446.c:236-252,520-603 uses; the "device" is a file writer (../stubs.c, ../README.md). */
#pragma once
#include <stdbool.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned long rtaudio_format_t;
#define RTAUDIO_FORMAT_FLOAT32 0x10
typedef unsigned int rtaudio_stream_flags_t;
#define RTAUDIO_FLAGS_NONINTERLEAVED 0x1
#define RTAUDIO_FLAGS_MINIMIZE_LATENCY 0x2
#define RTAUDIO_FLAGS_HOG_DEVICE 0x4
typedef unsigned int rtaudio_stream_status_t;
typedef enum rtaudio_error { RTAUDIO_ERROR_NONE = 0, RTAUDIO_ERROR_WARNING } rtaudio_error_t;
typedef enum rtaudio_api { RTAUDIO_API_UNSPECIFIED = 0, RTAUDIO_API_DUMMY, RTAUDIO_API_NUM } rtaudio_api_t;
typedef int (*rtaudio_cb_t)(void *out, void *in, unsigned int nFrames, double stream_time, rtaudio_stream_status_t status, void *userdata);
typedef void (*rtaudio_error_cb_t)(rtaudio_error_t err, const char *msg);
typedef struct rtaudio_device_info {
  unsigned int id, output_channels, input_channels, duplex_channels;
  int is_default_output, is_default_input;
  rtaudio_format_t native_formats;
  unsigned int preferred_sample_rate;
  unsigned int sample_rates[16];
  char name[512];
} rtaudio_device_info_t;
typedef struct rtaudio_stream_parameters { unsigned int device_id, num_channels, first_channel; } rtaudio_stream_parameters_t;
typedef struct rtaudio_stream_options { rtaudio_stream_flags_t flags; unsigned int num_buffers; int priority; char name[512]; } rtaudio_stream_options_t;
typedef struct rtaudio *rtaudio_t;
unsigned int rtaudio_get_num_compiled_apis(void);
const rtaudio_api_t *rtaudio_compiled_api(void);
const char *rtaudio_api_name(rtaudio_api_t api);
rtaudio_t rtaudio_create(rtaudio_api_t api);
void rtaudio_destroy(rtaudio_t audio);
rtaudio_api_t rtaudio_current_api(rtaudio_t audio);
int rtaudio_device_count(rtaudio_t audio);
unsigned int rtaudio_get_device_id(rtaudio_t audio, int i);
rtaudio_device_info_t rtaudio_get_device_info(rtaudio_t audio, unsigned int id);
unsigned int rtaudio_get_default_output_device(rtaudio_t audio);
void rtaudio_show_warnings(rtaudio_t audio, int show);
rtaudio_error_t rtaudio_open_stream(rtaudio_t audio, rtaudio_stream_parameters_t *output_params, rtaudio_stream_parameters_t *input_params,
                                    rtaudio_format_t format, unsigned int sample_rate, unsigned int *buffer_frames, rtaudio_cb_t cb,
                                    void *userdata, rtaudio_stream_options_t *options, rtaudio_error_cb_t errcb);
rtaudio_error_t rtaudio_start_stream(rtaudio_t audio);
rtaudio_error_t rtaudio_stop_stream(rtaudio_t audio);
int rtaudio_is_stream_open(rtaudio_t audio);
void rtaudio_close_stream(rtaudio_t audio);
#ifdef __cplusplus
}
#endif
