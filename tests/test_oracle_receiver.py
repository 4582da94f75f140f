"""CPU checks of the receiver oracle (oracle/receiver.c): RSSI, squelch / selector transitions, CTCSS detection.

The reference has no tests for this logic; these pin the restatement against independent
numpy computations and against the behaviour the reference's log lines describe
(/root/reference/src/sdr_pmr446.c:605-628, :828-874).
"""
import numpy as np

import rx_scenarios as sc
from oracle import oracle as orc

EV_TUNED, EV_CHANGED, EV_DETUNED, EV_ACQ, EV_CODE, EV_LOST = 1, 2, 4, 8, 16, 32
TONES = np.array([67.0, 71.9, 74.4, 77.0, 79.7, 82.5, 85.4, 88.5, 91.5, 94.8, 97.4, 100.0, 103.5, 107.2, 110.9, 114.8, 118.8, 123.0,
                  127.3, 131.8, 136.5, 141.3, 146.2, 151.4, 156.7, 162.2, 167.9, 173.8, 179.9, 186.2, 192.8, 203.5, 210.7, 218.1, 225.7,
                  233.6, 241.8, 250.3])


def _run(carriers, **kw):
    iq = sc.capture(carriers)
    o = orc.RxOracle(fs_in=sc.FS, in_fmt=1, chunk=sc.CHUNK, audio_gain=1.0, **kw)
    rows = o.run(iq, sc.CHUNK)
    o.close()
    return iq, rows


def test_rssi_is_mean_magnitude_in_db():
    car = sc.keyed_two_calls()
    iq, rows = _run(car)
    p = orc.PmrOracle(fs_in=sc.FS, in_fmt=1, chunk=sc.CHUNK)
    for k in range(3):
        c = p.execute(iq[2 * sc.CHUNK * k:2 * sc.CHUNK * (k + 1)], want=("chan",))["chan"]
        want = 20 * np.log10(np.mean(np.abs(c.astype(np.complex128)), axis=1))
        assert np.max(np.abs(rows[k]["rssi_ch"] - want)) < 1e-3
        assert abs(rows[k]["rssi"] - (want.max() - want.mean())) < 1e-3


def test_tune_detune_retune_and_ctcss_codes():
    car = sc.keyed_two_calls()
    _, rows = _run(car)
    act = [r["active_chan"] for r in rows]
    ev = [r["events"] for r in rows]
    assert act[0] == 1 and ev[0] & EV_TUNED                      # channel 2 opens the squelch in the first chunk
    k_det = next(k for k, e in enumerate(ev) if e & EV_DETUNED)
    assert 10 <= k_det <= 11 and act[k_det] == -1 and rows[k_det]["n_audio"] == 0   # carrier drops at 1.0 s = chunk 10.24
    assert rows[k_det]["ctcss_freq"] == 0.0 and not rows[k_det]["tone_detected"]
    k_re = next(k for k in range(k_det, len(rows)) if ev[k] & EV_TUNED)
    assert 15 <= k_re <= 16 and act[k_re] == 6                   # channel 7 keys up at 1.5 s = chunk 15.36
    assert all(a == -1 for a in act[k_det:k_re])
    # CTCSS: 67.0 Hz (code 1) acquired once 2441 samples went through, lost with the detune, then 88.5 Hz (code 8)
    k_acq = next(k for k, e in enumerate(ev) if e & EV_ACQ)
    assert k_acq == 1 and rows[k_acq]["ctcss_index"] == 0 and rows[k_acq]["ctcss_freq"] == np.float32(67.0)
    k_acq2 = next(k for k in range(k_re, len(rows)) if ev[k] & EV_ACQ)
    assert rows[k_acq2]["ctcss_index"] == 7 and rows[-1]["tone_detected"] and rows[-1]["ctcss_freq"] == np.float32(88.5)
    # hysteresis: every tuned chunk has margin > squelch - 5
    assert all(r["rssi"] >= 13.0 for r in rows if r["state"] == 1)


def test_lock_modes_and_channel_mask():
    car = sc.stronger_later()
    _, start = _run(car, lock_mode=0)
    _, follow = _run(car, lock_mode=1)
    assert all(r["active_chan"] == 7 for r in start)             # lock_mode_start stays where the squelch opened
    assert follow[0]["active_chan"] == 7 and follow[-1]["active_chan"] == 14
    k_ch = [k for k, r in enumerate(follow) if r["events"] & EV_CHANGED]
    assert len(k_ch) == 1 and 12 <= k_ch[0] <= 13                # 1.2 s = chunk 12.3
    assert follow[-1]["tone_detected"] and follow[-1]["ctcss_index"] == 37
    assert any(r["events"] & EV_CODE for r in follow[k_ch[0]:])  # 123.0 -> 250.3 Hz without losing the tone
    # masking channel 15 out keeps the receiver on channel 8 even when following the maximum
    _, masked = _run(car, lock_mode=1, channel_mask=(2 ** 64 - 1) & ~(1 << 14))
    assert all(r["active_chan"] == 7 for r in masked)


def test_goertzel_powers_against_float64():
    car = sc.keyed_two_calls()
    _, rows = _run(car)
    x = np.concatenate([r["ctcss_in"] for r in rows[:3]]).astype(np.float64)[:2441]
    coef = (2.0 * np.cos(np.float32(2.0 * np.pi * TONES.astype(np.float32).astype(np.float64) / 12500.0).astype(np.float64)))
    u0 = np.zeros(38)
    u1 = np.zeros(38)
    for v in x:
        u0, u1 = v + coef * u0 - u1, u0
    want = u0 * u0 + u1 * u1 - coef * u0 * u1
    got = rows[2]["ctcss_power"]
    assert np.argmax(got) == np.argmax(want) == 0
    assert np.max(np.abs(got - want)) / want.max() < 2e-3
