#!/usr/bin/env python3
"""Device-resident throughput of the BASELINE.json configurations that are NOT the bench.py line (they are parity-test
cases; this probe only reports how fast each runs).  Prints one JSON object; not a judged number."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ctypes as C

from sdr_pmr446_b200 import _lib, chain, synth


def timed(fn, steps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def tiled(base, S):
    t = torch.from_numpy(base).cuda()
    return t.unsqueeze(0).repeat(S, *([1] * t.dim())).contiguous()


out = {}
# configs[0] scaled out: 1.024 Msps cu8, reference chunk 100 000, 16 channels -> s16 (1024 streams)
S, fs, n = 1024, 1024000, 1000000
iq = tiled(synth.make_cu8(synth.CaptureSpec(fs=float(fs)), n, 446), S)
b = chain.PmrBatch(n_streams=S, fs_in=fs, in_fmt=1, audio_gain=1.0, max_chunk=n)
pcm = torch.empty((S, 16, b.max_ns), dtype=torch.int16, device="cuda")
ms = timed(lambda: b.execute_device(iq, n, {"pcm": pcm, "ld": b.max_ns}))
out["cfg0_1024k_cu8_x1024"] = {"ms_per_step": ms, "msps": S * n / ms / 1e3, "note": "1 s... %d samples per stream per step" % n}
b.close()
del iq, pcm
# configs[1]: dsd_in chain, 2.4 Msps cu8 -> 48 kHz s16 (1024 streams)
S, fs, n = 1024, 2400000, 1200000
iq = tiled(synth.cfg2_capture(n=n), S)
d = chain.DsdBatch(n_streams=S, fs_in=fs, in_fmt=1, max_chunk=n)
pcm = torch.empty((S, d.max_out), dtype=torch.int16, device="cuda")
ms = timed(lambda: d.execute_device(iq, n, pcm=pcm))
out["cfg1_dsd_2400k_cu8_x1024"] = {"ms_per_step": ms, "msps": S * n / ms / 1e3}
d.close()
del iq, pcm
# configs[3]: 20 Msps cf32, 1600 channels + waterfall W = 1600 (4 streams, 0.1 s per step)
S, fs, n, M = 4, 20000000, 2000000, 1600
car = tuple(synth.Carrier(int(c), 0.05, 1000.0, 67.0) for c in np.random.default_rng(446).choice(np.arange(1, M + 1), 64, replace=False))
iq = tiled(synth.make_cf32(synth.CaptureSpec(fs=float(fs), num_channels=M, carriers=car), n, 446), S)
b = chain.PmrBatch(n_streams=S, fs_in=fs, in_fmt=0, num_channels=M, waterfall=1600, audio_gain=1.0, max_chunk=n)
pcm = torch.empty((S, M, b.max_ns), dtype=torch.int16, device="cuda")
asc = torch.empty((S, 1600), dtype=torch.uint8, device="cuda")
pk = torch.empty((S, 2), dtype=torch.float32, device="cuda")
ms = timed(lambda: b.execute_device(iq.view(torch.float32).view(S, -1), n, {"pcm": pcm, "ld": b.max_ns, "ascii": asc, "peak": pk}))
out["cfg3_wideband_20M_cf32_1600ch_x4"] = {"ms_per_step": ms, "msps": S * n / ms / 1e3, "realtime_streams": S * n / ms / 1e3 / 20.0}
b.close()
# receiver mode (what the reference plays): 1024 x 2.4 Msps, RSSI + squelch + selected-channel audio + CTCSS per stream
S, fs, n = 1024, 2400000, 2400000
iq = tiled(synth.make_cu8(synth.CaptureSpec(fs=float(fs)), n, 446), S)
rx = chain.PmrReceiver(n_streams=S, fs_in=fs, in_fmt=1, max_chunk=n, audio_gain=1.0)
o = {"ld": rx.max_ns, "pcm": torch.empty((S, rx.max_ns), dtype=torch.int16, device="cuda"),
     "status": torch.empty((S, C.sizeof(_lib.RxStatus)), dtype=torch.uint8, device="cuda"),
     "rssi": torch.empty((S, 16), dtype=torch.float32, device="cuda")}
ms = timed(lambda: rx.execute_device(iq, n, o))
out["receiver_2400k_cu8_x1024"] = {"ms_per_step": ms, "msps": S * n / ms / 1e3, "launches": rx.last_launches}
rx.close()
print(json.dumps(out, indent=1))
