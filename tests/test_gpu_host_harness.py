"""The C host harnesses (host/*.c, derived from the reference's main() loops) on the GPU box: the same liquid-API
source linked against libpmr446_b200.so and against the CPU oracle must write the same s16 audio; the coarse
batched harness must agree with the oracle for every channel."""
import os
import subprocess

import numpy as np
import pytest

from util import PCM_TOL_LSB

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])


def test_liquid_loop_gpu_equals_cpu(tmp_path):
    from sdr_pmr446_b200 import synth
    _build()
    n = 300000
    x = synth.make_cf32(synth.CaptureSpec(fs=1024000.0), n, 446)
    cap = tmp_path / "cap.cf32"
    x.tofile(cap)
    outs = {}
    for kind in ("gpu", "cpu"):
        o = tmp_path / ("audio_%s.s16" % kind)
        subprocess.check_call([os.path.join(ROOT, "host", "pmr446_liquid_loop_" + kind), "-i", str(cap), "-o", str(o), "-c", "8"],
                              stdout=subprocess.DEVNULL)
        outs[kind] = np.fromfile(o, np.int16)
    assert outs["gpu"].size == outs["cpu"].size == 3662
    d = np.abs(outs["gpu"][500:].astype(np.int32) - outs["cpu"][500:].astype(np.int32))
    assert d.max() <= PCM_TOL_LSB


def test_batch_file_harness_all_channels(tmp_path):
    from oracle import oracle as orc
    from sdr_pmr446_b200 import synth
    _build()
    n, fs = 500000, 2400000
    caps = []
    for s in range(2):
        iq = synth.make_cu8(synth.CaptureSpec(fs=float(fs), carriers=synth.rotated_carriers(s)), n, 446 + s)
        p = tmp_path / ("cap%d.cu8" % s)
        iq.tofile(p)
        caps.append((p, iq))
    prefix = tmp_path / "out"
    subprocess.check_call([os.path.join(ROOT, "host", "pmr446_batch_file"), "-r", str(fs), "-8", "-n", "240000", "-o", str(prefix)] +
                          [str(p) for p, _ in caps])
    for s, (_, iq) in enumerate(caps):
        o = orc.PmrOracle(fs_in=fs, in_fmt=1, audio_gain=1.0, chunk=240000)
        r = o.run(iq, 240000)
        o.close()
        for c in sorted({x.channel - 1 for x in synth.rotated_carriers(s)}):
            g = np.fromfile("%s.%d.%02d.s16" % (prefix, s, c + 1), np.int16)
            assert g.size == r["pcm"].shape[1]
            d = np.abs(g[500:].astype(np.int32) - r["pcm"][c, 500:].astype(np.int32))
            assert d.max() <= PCM_TOL_LSB, (s, c, int(d.max()))


def test_rx_file_harness_wav_and_log(tmp_path):
    """pmr446_rx_file: the reference program on a capture file -- WAV of the selected channel, the reference's log lines."""
    import struct

    import rx_scenarios as sc
    from oracle import oracle as orc
    _build()
    car = sc.keyed_two_calls()
    iq = sc.capture(car, seconds=2.5)
    cap = tmp_path / "cap.cu8"
    iq.tofile(cap)
    wav = tmp_path / "out.wav"
    p = subprocess.run([os.path.join(ROOT, "host", "pmr446_rx_file"), "-8", "-g", "1.0", "-o", str(wav), str(cap)], check=True,
                       capture_output=True, text=True)
    log = p.stderr
    assert "Tuned to channel 2" in log and "Detuned from channel 2" in log and "Tuned to channel 7" in log
    assert "Acquired CTCSS code: 1 (frequency: 67.00Hz)" in log and "Acquired CTCSS code: 8 (frequency: 88.50Hz)" in log
    raw = wav.read_bytes()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[36:40] == b"data"
    fmt, ch, rate, _, _, bits = struct.unpack("<HHIIHH", raw[20:36])
    assert (fmt, ch, rate, bits) == (1, 1, 12500, 16)
    data_len = struct.unpack("<I", raw[40:44])[0]
    pcm = np.frombuffer(raw[44:], np.int16)
    assert data_len == pcm.size * 2 and struct.unpack("<I", raw[4:8])[0] == 36 + data_len
    o = orc.RxOracle(fs_in=sc.FS, in_fmt=1, chunk=sc.CHUNK, audio_gain=1.0)
    rows = o.run(iq, sc.CHUNK)
    o.close()
    ref = np.concatenate([r["pcm"] for r in rows])
    assert pcm.size == ref.size
    # strict on the steady part of the first call (chunks 2..9), see test_gpu_receiver_parity for why
    a, b = 2 * 1221, 9 * 1220
    assert np.abs(pcm[a:b].astype(np.int32) - ref[a:b].astype(np.int32)).max() <= PCM_TOL_LSB


def test_rx_file_harness_waterfall_footer(tmp_path):
    import rx_scenarios as sc
    _build()
    iq = sc.capture(sc.stronger_later(), seconds=0.6)
    cap = tmp_path / "cap.cu8"
    iq.tofile(cap)
    p = subprocess.run([os.path.join(ROOT, "host", "pmr446_rx_file"), "-8", "-w", "64", "-m", "3,4", "-f", "-o", str(tmp_path / "o.wav"), str(cap)],
                       check=True, capture_output=True, text=True)
    rows = [ln for ln in p.stdout.split("\n") if ln.startswith(" > ")]
    assert len(rows) == 7 and all(len(r.split(" < ")[0]) == 3 + 64 for r in rows)
    foot = [ln for ln in p.stdout.split("\n") if "MHz" in ln][-1]   # text mode turns the footer's \r into \n
    assert "^^" in foot and "--" in foot and "446.100 MHz [8]" in foot and " 01 " in foot
    assert (tmp_path / "o.wav").read_bytes()[20:22] == b"\x03\x00"   # IEEE float WAV


def test_dsd_pipe_harness(tmp_path):
    """dsd446_pipe writes the dsd_in s16 stream to stdout (what `| dsd -i -` reads)."""
    from oracle import oracle as orc
    from sdr_pmr446_b200 import synth
    _build()
    iq = synth.cfg2_capture(n=1_200_000)
    cap = tmp_path / "cap.cu8"
    iq.tofile(cap)
    with open(cap, "rb") as fi:
        p = subprocess.run([os.path.join(ROOT, "host", "dsd446_pipe"), "-r", "2400000", "-8", "-"], stdin=fi, check=True, capture_output=True)
    pcm = np.frombuffer(p.stdout, np.int16)
    o = orc.DsdOracle(fs_in=2400000, in_fmt=1, chunk=200000)
    r = o.run(iq, 200000)
    assert pcm.size == r["pcm"].size == r["nz"]
    assert np.abs(pcm[200:].astype(np.int32) - r["pcm"][200:].astype(np.int32)).max() <= PCM_TOL_LSB
