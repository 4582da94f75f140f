"""Pins oracle/ against the REAL liquid-dsp (v1.7.0, /root/reference/.github/workflows/build.yml:30) the moment a box has
it: every object of the restatement is driven side by side with libliquid's through the same ctypes prototypes
(tests/liquid_api.py) and every stage is diffed -- sample counts equal, float stages within 1e-4 relative RMS, waterfall
rows within 0.02 dB.  liquid-dsp is NOT in this image (no liquid.h, no libliquid, no network), so here these tests SKIP;
`LIQUID_LIB=/path/libliquid.so pytest tests/test_oracle_vs_liquid.py` runs them, and
`make -f oracle/ref.mk LIQUID=system` + tests/test_reference_build.py runs the unmodified reference programs on it.
The Appendix-A items flagged "verify first" (resampler fc / npfb, Kaiser r, half-band design, asgram column rule) are
exactly what a failure here would point at; they are runtime knobs of the oracle (oracle_liquid_get_knobs)."""
import ctypes as C
import ctypes.util
import os

import numpy as np
import pytest

import liquid_api
from util import REL_RMS_TOL, rel_rms


def _real_liquid():
    path = os.environ.get("LIQUID_LIB") or ctypes.util.find_library("liquid")
    if not path:
        pytest.skip("liquid-dsp is not installed (set LIQUID_LIB=/path/to/libliquid.so)")
    return liquid_api.bind(C.CDLL(path))


def _oracle():
    from oracle import oracle as orc
    return liquid_api.bind(orc.lib())


def _capture(fs, n, seed=446):
    from sdr_pmr446_b200 import synth
    x = synth.make_cu8(synth.CaptureSpec(fs=float(fs)), n, seed).astype(np.float32)
    x = (x - np.float32(127.4)) * np.float32(1.0 / 128.0)
    return (x[0::2] + 1j * x[1::2]).astype(np.complex64)


def test_reference_loop_every_stage():
    """init_liquid() + the main-loop body (src/sdr_pmr446.c:420-480, :795-913) on both libraries."""
    R, O = _real_liquid(), _oracle()
    hp, lp = liquid_api.reference_taps()
    iq = _capture(1024000, 400000)
    r = liquid_api.reference_loop(R, iq, hp, lp, active_chan=1, lowpass=True, waterfall=120)
    o = liquid_api.reference_loop(O, iq, hp, lp, active_chan=1, lowpass=True, waterfall=120)
    assert o["res"].size == r["res"].size and o["chan"].shape == r["chan"].shape
    assert rel_rms(o["res"], r["res"]) < REL_RMS_TOL
    assert rel_rms(o["chan"], r["chan"]) < REL_RMS_TOL
    for c in range(16):
        assert rel_rms(o["chan"][c], r["chan"][c]) < 5e-4, c
    assert rel_rms(o["audio"][500:], r["audio"][500:]) < REL_RMS_TOL
    for (oa, op, of), (ra, rp, rf) in zip(o["rows"], r["rows"]):
        assert abs(op - rp) < 0.02 and abs(of - rf) < 1e-6
        assert np.sum(oa != ra) <= 2


@pytest.mark.parametrize("rate", [200000 / 1024000, 200000 / 2400000, 12500 / 1024000, 12500 / 2400000, 200000 / 3200000, 0.625])
def test_msresamp_crcf_rates(rate):
    """msresamp_crcf_create(rate, 60) for every plan the configs use (src/sdr_pmr446.c:425-426, src/dsd_in.c:100)."""
    R, O = _real_liquid(), _oracle()
    x = _capture(1024000, 300000)
    outs = []
    for L in (R, O):
        q = L.msresamp_crcf_create(np.float32(rate), 60.0)
        y = np.zeros(x.size + 64, np.complex64)
        tot = []
        for o in range(0, x.size, 100000):
            xb = x[o:o + 100000].copy()
            ny = C.c_uint()
            assert L.msresamp_crcf_execute(q, xb.ctypes.data, xb.size, y.ctypes.data, C.byref(ny)) == 0
            tot.append(y[:ny.value].copy())
        L.msresamp_crcf_destroy(q)
        outs.append(np.concatenate(tot))
    assert outs[0].size == outs[1].size
    assert rel_rms(outs[1], outs[0]) < REL_RMS_TOL


def test_msresamp_rrrf_interpolator():
    """msresamp_rrrf_create(48000 / 12500, 60) (src/dsd_in.c:104)."""
    R, O = _real_liquid(), _oracle()
    rng = np.random.default_rng(446)
    x = (0.3 * np.sin(2 * np.pi * 1000.0 / 12500.0 * np.arange(20000)) + 0.01 * rng.standard_normal(20000)).astype(np.float32)
    outs = []
    for L in (R, O):
        q = L.msresamp_rrrf_create(np.float32(48000.0) / np.float32(12500.0), 60.0)
        y = np.zeros(4 * x.size + 64, np.float32)
        nz = C.c_uint()
        assert L.msresamp_rrrf_execute(q, x.ctypes.data, x.size, y.ctypes.data, C.byref(nz)) == 0
        L.msresamp_rrrf_destroy(q)
        outs.append(y[:nz.value].copy())
    assert outs[0].size == outs[1].size
    assert rel_rms(outs[1], outs[0]) < REL_RMS_TOL


@pytest.mark.parametrize("W", [64, 120, 250, 1600])
def test_asgram_rows(W):
    """asgramcf_create(W) / set_scale(-40, 2) / write / execute (src/sdr_pmr446.c:473-477, :910-913)."""
    R, O = _real_liquid(), _oracle()
    x = _capture(1024000, 60000)[:39000]
    rows = []
    for L in (R, O):
        q = L.asgramcf_create(W)
        L.asgramcf_set_scale(q, -40.0, 2.0)
        got = []
        for k in range(3):
            xb = x[k * 13000:(k + 1) * 13000].copy()
            row = np.zeros(W, np.uint8)
            pv, pf = C.c_float(), C.c_float()
            assert L.asgramcf_write(q, xb.ctypes.data, xb.size) == 0
            assert L.asgramcf_execute(q, row.ctypes.data, C.byref(pv), C.byref(pf)) == 0
            got.append((row, pv.value, pf.value))
        L.asgramcf_destroy(q)
        rows.append(got)
    for (ra, rp, rf), (oa, op, of) in zip(*rows):
        assert abs(op - rp) < 0.02 and abs(of - rf) < 1e-6
        assert np.sum(oa != ra) <= max(2, W // 100)
