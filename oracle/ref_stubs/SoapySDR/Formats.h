/* TEST INFRASTRUCTURE: stand-in for <SoapySDR/Formats.h> (see ../README.md). */
#pragma once
#define SOAPY_SDR_CF32 "CF32"
