"""ctypes prototypes of the liquid-dsp subset (include/pmr446_liquid_shim.h == oracle/liquid_subset.h), applicable
to either shared library, plus a transcription of the reference's loop body on top of them."""
import ctypes as C

import numpy as np

NUM_CHANNELS = 16


def bind(L):
    vp, u, f, i = C.c_void_p, C.c_uint, C.c_float, C.c_int
    P = C.POINTER
    sig = {
        "iirfilt_crcf_create_dc_blocker": (vp, [f]), "iirfilt_crcf_execute_block": (i, [vp, vp, u, vp]), "iirfilt_crcf_destroy": (i, [vp]),
        "iirfilt_rrrf_create": (vp, [vp, u, vp, u]), "iirfilt_rrrf_create_dc_blocker": (vp, [f]),
        "iirfilt_rrrf_execute_block": (i, [vp, vp, u, vp]), "iirfilt_rrrf_destroy": (i, [vp]),
        "msresamp_crcf_create": (vp, [f, f]), "msresamp_crcf_execute": (i, [vp, vp, u, vp, P(u)]), "msresamp_crcf_destroy": (i, [vp]),
        "msresamp_rrrf_create": (vp, [f, f]), "msresamp_rrrf_execute": (i, [vp, vp, u, vp, P(u)]), "msresamp_rrrf_destroy": (i, [vp]),
        "nco_crcf_create": (vp, [i]), "nco_crcf_set_frequency": (i, [vp, f]), "nco_crcf_step": (i, [vp]),
        "nco_crcf_mix_block_down": (i, [vp, vp, vp, u]), "nco_crcf_destroy": (i, [vp]),
        "firpfbch_crcf_create_kaiser": (vp, [i, u, u, f]), "firpfbch_crcf_analyzer_execute": (i, [vp, vp, vp]),
        "firpfbch_crcf_destroy": (i, [vp]),
        "freqdem_create": (vp, [f]), "freqdem_demodulate_block": (i, [vp, vp, u, vp]), "freqdem_reset": (i, [vp]), "freqdem_destroy": (i, [vp]),
        "firfilt_rrrf_create": (vp, [vp, u]), "firfilt_rrrf_execute_block": (i, [vp, vp, u, vp]), "firfilt_rrrf_destroy": (i, [vp]),
        "wdelayf_create": (vp, [u]), "wdelayf_push": (i, [vp, f]), "wdelayf_read": (i, [vp, P(f)]), "wdelayf_destroy": (i, [vp]),
        "cbuffercf_create": (vp, [u]), "cbuffercf_write": (i, [vp, vp, u]), "cbuffercf_size": (u, [vp]),
        "cbuffercf_read": (i, [vp, u, P(vp), P(u)]), "cbuffercf_release": (i, [vp, u]), "cbuffercf_destroy": (i, [vp]),
        "asgramcf_create": (vp, [u]), "asgramcf_set_scale": (i, [vp, f, f]), "asgramcf_write": (i, [vp, vp, u]),
        "asgramcf_execute": (i, [vp, vp, P(f), P(f)]), "asgramcf_destroy": (i, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L


SHIM_SYMBOLS = [
    "iirfilt_crcf_create_dc_blocker", "iirfilt_crcf_execute_block", "iirfilt_crcf_destroy", "iirfilt_rrrf_create",
    "iirfilt_rrrf_create_dc_blocker", "iirfilt_rrrf_execute_block", "iirfilt_rrrf_destroy", "msresamp_crcf_create",
    "msresamp_crcf_execute", "msresamp_crcf_print", "msresamp_crcf_destroy", "msresamp_rrrf_create", "msresamp_rrrf_execute",
    "msresamp_rrrf_print", "msresamp_rrrf_destroy", "nco_crcf_create", "nco_crcf_set_frequency", "nco_crcf_mix_down", "nco_crcf_step",
    "nco_crcf_mix_block_down", "nco_crcf_destroy", "firpfbch_crcf_create_kaiser", "firpfbch_crcf_analyzer_execute",
    "firpfbch_crcf_destroy", "freqdem_create", "freqdem_demodulate_block", "freqdem_reset", "freqdem_destroy", "firfilt_rrrf_create",
    "firfilt_rrrf_execute_block", "firfilt_rrrf_destroy", "wdelayf_create", "wdelayf_push", "wdelayf_read", "wdelayf_destroy",
    "cbuffercf_create", "cbuffercf_write", "cbuffercf_size", "cbuffercf_read", "cbuffercf_release", "cbuffercf_destroy",
    "cbufferf_create", "cbufferf_write", "cbufferf_size", "cbufferf_max_size", "cbufferf_read", "cbufferf_release", "cbufferf_destroy",
    "asgramcf_create", "asgramcf_set_scale", "asgramcf_write", "asgramcf_execute", "asgramcf_destroy",
]


def reference_loop(L, iq_cf32, hp, lp, active_chan, chunk=100000, audio_gain=1.0, lowpass=False, waterfall=0, fs=1024000):
    """init_liquid() + main-loop body of src/sdr_pmr446.c:420-480,795-913 for one fixed active channel."""
    p = lambda a: a.ctypes.data
    dcblock = L.iirfilt_crcf_create_dc_blocker(0.0005)
    resampler = L.msresamp_crcf_create(np.float32(200000.0) / np.float32(fs), 60.0)
    nco = L.nco_crcf_create(1)
    L.nco_crcf_set_frequency(nco, np.float32(-0.5 * 15 / 16 * 2 * np.pi))
    channelizer = L.firpfbch_crcf_create_kaiser(0, NUM_CHANNELS, 13, 80.0)
    fm_demod = L.freqdem_create(0.5)
    ctcss_filt = L.firfilt_rrrf_create(p(hp), hp.size)
    delay = L.wdelayf_create((hp.size - 1) // 2)
    audio_filt = L.firfilt_rrrf_create(p(lp), lp.size)
    b = np.array([0.507301437230636, 0.507301437230636], np.float32)
    a = np.array([1.0, 0.014602874461272194], np.float32)
    deemph = L.iirfilt_rrrf_create(p(b), 2, p(a), 2)
    ring = L.cbuffercf_create(39064)
    asgram = None
    if waterfall:
        asgram = L.asgramcf_create(waterfall)
        L.asgramcf_set_scale(asgram, -40.0, 2.0)
    assert all([dcblock, resampler, nco, channelizer, fm_demod, ctcss_filt, delay, audio_filt, deemph, ring])
    res_all, chan_all, audio_all, rows = [], [], [], []
    for o in range(0, iq_cf32.size, chunk):
        buffp = iq_cf32[o:o + chunk].copy()   # the DC blocker runs in place (:795): never touch the caller's capture
        n = buffp.size
        resamp_buf = np.zeros(39064, np.complex64)
        ny = C.c_uint()
        assert L.iirfilt_crcf_execute_block(dcblock, p(buffp), n, p(buffp)) == 0
        assert L.msresamp_crcf_execute(resampler, p(buffp), n, p(resamp_buf), C.byref(ny)) == 0
        res_all.append(resamp_buf[:ny.value].copy())
        assert L.cbuffercf_write(ring, p(resamp_buf), ny.value) == 0
        chan_bufs = np.zeros((NUM_CHANNELS, 2441), np.complex64)
        tmp_out = np.zeros(NUM_CHANNELS, np.complex64)
        ns = 0
        rpc, nr = C.c_void_p(), C.c_uint()
        while L.cbuffercf_size(ring) >= NUM_CHANNELS:
            L.cbuffercf_read(ring, NUM_CHANNELS, C.byref(rpc), C.byref(nr))
            L.nco_crcf_mix_block_down(nco, rpc, rpc, NUM_CHANNELS)
            L.firpfbch_crcf_analyzer_execute(channelizer, rpc, p(tmp_out))
            L.cbuffercf_release(ring, nr.value)
            chan_bufs[:, ns] = tmp_out
            ns += 1
        chan_all.append(chan_bufs[:, :ns].copy())
        tmp1 = np.zeros(2441, np.float32)
        tmp2 = np.zeros(2441, np.float32)
        cb = np.ascontiguousarray(chan_bufs[active_chan])
        L.freqdem_demodulate_block(fm_demod, p(cb), ns, p(tmp1))
        L.firfilt_rrrf_execute_block(ctcss_filt, p(tmp1), ns, p(tmp2))
        t = C.c_float()
        for k in range(ns):
            L.wdelayf_push(delay, float(tmp1[k]))
            L.wdelayf_read(delay, C.byref(t))
            tmp1[k] = np.float32(t.value) - tmp2[k]
            tmp2[k] *= np.float32(audio_gain)
        L.iirfilt_rrrf_execute_block(deemph, p(tmp2), ns, p(tmp2))
        if lowpass:
            L.firfilt_rrrf_execute_block(audio_filt, p(tmp2), ns, p(tmp2))
        audio_all.append(tmp2[:ns].copy())
        if asgram:
            row = np.zeros(waterfall, np.uint8)
            pv, pf = C.c_float(), C.c_float()
            L.asgramcf_write(asgram, p(resamp_buf), ny.value)
            L.asgramcf_execute(asgram, p(row), C.byref(pv), C.byref(pf))
            rows.append((row, pv.value, pf.value))
    for fn, h in (("iirfilt_crcf_destroy", dcblock), ("msresamp_crcf_destroy", resampler), ("nco_crcf_destroy", nco),
                  ("firpfbch_crcf_destroy", channelizer), ("freqdem_destroy", fm_demod), ("firfilt_rrrf_destroy", ctcss_filt),
                  ("wdelayf_destroy", delay), ("firfilt_rrrf_destroy", audio_filt), ("iirfilt_rrrf_destroy", deemph),
                  ("cbuffercf_destroy", ring)):
        getattr(L, fn)(h)
    if asgram:
        L.asgramcf_destroy(asgram)
    return {"res": np.concatenate(res_all), "chan": np.concatenate(chan_all, axis=1), "audio": np.concatenate(audio_all), "rows": rows}


def reference_taps():
    import os
    import re
    txt = open(os.path.join(os.path.dirname(__file__), "..", "include", "pmr446_taps.h")).read()

    def half(name, n):
        m = re.search(r"%s_half\[%d\] = \{(.*?)\};" % (name, n), txt, re.S)
        h = np.array([float(v.rstrip("f")) for v in re.findall(r"[-+]?\d+\.\d+f", m.group(1))], np.float32)
        assert h.size == n
        return np.ascontiguousarray(np.concatenate([h, h[-2::-1]]))
    return half("pmr446_hp_audio_taps", 189), half("pmr446_lp_audio_taps", 52)
