# TEST INFRASTRUCTURE.  Builds the UNMODIFIED reference programs from the sources where they lie under $(REF) into
# oracle/_ref/ (git-ignored; travels to the GPU box like any built file), with file-backed stand-ins for SoapySDR /
# RtAudio / dlg (oracle/ref_stubs/).  Run from the repo root:
#
#   make -f oracle/ref.mk                 LIQUID=oracle: <liquid/liquid.h> is the oracle's restatement (no liquid-dsp in
#                                         this image) -> pins oracle/chains.c and oracle/receiver.c against the reference's
#                                         own main loops, selector and CTCSS detector
#   make -f oracle/ref.mk LIQUID=system   the real liquid-dsp v1.7.0 (<liquid/liquid.h> + -lliquid on the default paths, or
#                                         LIQUID_PREFIX=/path) -> the true reference; tests/test_oracle_vs_liquid.py then diffs
#                                         every stage of oracle/ against it
#
# Nothing is copied from $(REF); no file of the reference's build system is run.
REF ?= /root/reference
CC ?= gcc
LIQUID ?= $(shell echo '#include <liquid/liquid.h>' | $(CC) $(if $(LIQUID_PREFIX),-I$(LIQUID_PREFIX)/include) -E -x c - >/dev/null 2>&1 && echo system || echo oracle)
OUT = oracle/_ref
STUBS = oracle/ref_stubs
CFLAGS = -O2 -g -std=gnu11 -ffp-contract=off -fno-stack-protector -I$(STUBS) -I$(REF)/include
ifeq ($(LIQUID),system)
  CFLAGS += $(if $(LIQUID_PREFIX),-I$(LIQUID_PREFIX)/include)
  LIBS = $(if $(LIQUID_PREFIX),-L$(LIQUID_PREFIX)/lib -Wl$(comma)-rpath$(comma)$(LIQUID_PREFIX)/lib) -lliquid -lm -lpthread
  DSP =
else
  CFLAGS += -I$(STUBS)/liquid_oracle
  LIBS = -lm -lpthread
  DSP = oracle/liquid_subset.c
endif
comma := ,

all: $(OUT)/sdr_pmr446_ref $(OUT)/dsd_in_ref $(OUT)/LIQUID

$(OUT)/LIQUID:
	@mkdir -p $(OUT)
	@echo $(LIQUID) > $@

# src/logging.c needs dlg's internals (dlg/output.h) and is the one reference file left out: stubs.c has logging_init()
$(OUT)/sdr_pmr446_ref: $(REF)/src/sdr_pmr446.c $(REF)/src/shared.c $(STUBS)/stubs.c $(DSP)
	@mkdir -p $(OUT)
	$(CC) $(CFLAGS) -DAPP_SDR_PMR446 -o $@ $^ $(LIBS)

$(OUT)/dsd_in_ref: $(REF)/src/dsd_in.c $(REF)/src/shared.c $(STUBS)/stubs.c $(DSP)
	@mkdir -p $(OUT)
	$(CC) $(CFLAGS) -DAPP_DSD_IN -o $@ $^ $(LIBS)

clean:
	rm -rf $(OUT)
.PHONY: all clean
