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
