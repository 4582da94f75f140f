/* TEST INFRASTRUCTURE (LIQUID=oracle builds of oracle/ref.mk): <liquid/liquid.h> forwarded to the CPU restatement of
 * the liquid-dsp subset the reference calls, so the reference's unmodified main loops run over oracle/liquid_subset.c. */
#pragma once
#include "../../../liquid_subset.h"
