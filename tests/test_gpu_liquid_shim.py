"""GPU tests of the liquid-signature shim: the reference's loop body, transcribed once (liquid_api.reference_loop),
runs against libpmr446_b200.so (every block call executes on the GPU) and against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import liquid_api
from util import REL_RMS_TOL, rel_rms

pytestmark = pytest.mark.gpu


def _libs():
    from oracle import oracle as orc
    from sdr_pmr446_b200 import _lib
    G = liquid_api.bind(C.CDLL(_lib.LIB_PATH))
    O = liquid_api.bind(orc.lib())
    return G, O


def test_reference_loop_body_on_the_shim():
    from sdr_pmr446_b200 import synth
    G, O = _libs()
    hp, lp = liquid_api.reference_taps()
    n = 250000
    x = synth.make_cu8(synth.CaptureSpec(fs=1024000.0), n, 446).astype(np.float32)
    x = ((x - np.float32(127.4)) * np.float32(1.0 / 128.0))
    iq = (x[0::2] + 1j * x[1::2]).astype(np.complex64)
    g = liquid_api.reference_loop(G, iq, hp, lp, active_chan=1, lowpass=True, waterfall=120)
    r = liquid_api.reference_loop(O, iq, hp, lp, active_chan=1, lowpass=True, waterfall=120)
    assert g["res"].size == r["res"].size and g["chan"].shape == r["chan"].shape
    assert rel_rms(g["res"], r["res"]) < REL_RMS_TOL
    assert rel_rms(g["chan"], r["chan"]) < REL_RMS_TOL
    assert rel_rms(g["chan"][1], r["chan"][1]) < REL_RMS_TOL
    k = 1 if g["audio"][0] == r["audio"][0] else 500
    assert rel_rms(g["audio"][k:], r["audio"][k:]) < REL_RMS_TOL
    for (ga, gp, gf), (ra, rp, rf) in zip(g["rows"], r["rows"]):
        assert abs(gp - rp) < 0.02 and abs(gf - rf) < 1e-6
        assert np.sum(ga != ra) <= 2


def test_upsampler_and_small_objects():
    G, O = _libs()
    rng = np.random.default_rng(7)
    # msresamp_rrrf x3.84 (src/dsd_in.c:104,170), fed in uneven blocks
    x = (0.4 * np.sin(2 * np.pi * 1000.0 / 12500.0 * np.arange(5000)) + 0.01 * rng.standard_normal(5000)).astype(np.float32)
    outs = []
    for L in (G, O):
        q = L.msresamp_rrrf_create(48000.0 / 12500.0, 60.0)
        assert q
        ys = []
        for a, b in ((0, 1), (1, 1234), (1234, 5000)):
            blk = np.ascontiguousarray(x[a:b])
            y = np.zeros(4 * blk.size + 8, np.float32)
            ny = C.c_uint()
            assert L.msresamp_rrrf_execute(q, blk.ctypes.data, blk.size, y.ctypes.data, C.byref(ny)) == 0
            ys.append(y[:ny.value])
        L.msresamp_rrrf_destroy(q)
        outs.append(np.concatenate(ys))
    assert outs[0].size == outs[1].size and rel_rms(outs[0], outs[1]) < REL_RMS_TOL
    # real DC blocker on the CTCSS branch (src/sdr_pmr446.c:450,606) and freqdem_reset (:866)
    y = []
    for L in (G, O):
        q = L.iirfilt_rrrf_create_dc_blocker(0.0005)
        v = (x + 0.3).astype(np.float32)
        o = np.zeros_like(v)
        L.iirfilt_rrrf_execute_block(q, v.ctypes.data, 3000, o.ctypes.data)
        L.iirfilt_rrrf_execute_block(q, v[3000:].ctypes.data, 2000, o[3000:].ctypes.data)
        L.iirfilt_rrrf_destroy(q)
        y.append(o)
    assert rel_rms(y[0], y[1]) < REL_RMS_TOL
    z = (rng.standard_normal(64) + 1j * rng.standard_normal(64)).astype(np.complex64)
    m = []
    for L in (G, O):
        q = L.freqdem_create(0.5)
        o = np.zeros(64, np.float32)
        L.freqdem_demodulate_block(q, z.ctypes.data, 32, o.ctypes.data)
        L.freqdem_reset(q)
        L.freqdem_demodulate_block(q, z[32:].ctypes.data, 32, o[32:].ctypes.data)
        L.freqdem_destroy(q)
        m.append(o)
    assert m[0][32] == m[1][32] and np.max(np.abs(m[0][1:] - m[1][1:])) < 2e-6
