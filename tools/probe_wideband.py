#!/usr/bin/env python3
"""Per-kernel timing of the 1600-channel wideband configuration (development probe)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sdr_pmr446_b200 import chain, synth

S, fs, n, M = int(os.environ.get("WB_STREAMS", "4")), 20000000, 2000000, 1600
W = int(os.environ.get("WB_WATERFALL", "1600"))
car = tuple(synth.Carrier(int(c), 0.05, 1000.0, 67.0) for c in np.random.default_rng(446).choice(np.arange(1, M + 1), 64, replace=False))
base = torch.from_numpy(synth.make_cf32(synth.CaptureSpec(fs=float(fs), num_channels=M, carriers=car), n, 446)).cuda()
iq = base.unsqueeze(0).repeat(S, 1).contiguous()
b = chain.PmrBatch(n_streams=S, fs_in=fs, in_fmt=0, num_channels=M, waterfall=W, audio_gain=1.0, max_chunk=n)
outs = {"pcm": torch.empty((S, M, b.max_ns), dtype=torch.int16, device="cuda"), "ld": b.max_ns}
if W:
    outs["ascii"] = torch.empty((S, W), dtype=torch.uint8, device="cuda")
    outs["peak"] = torch.empty((S, 2), dtype=torch.float32, device="cuda")
x = iq.view(torch.float32).view(S, -1)
for _ in range(2):
    b.execute_device(x, n, outs)
b.timing(True)
for _ in range(3):
    b.execute_device(x, n, outs)
tm = b.get_timings()
print({k: round(v[0] / v[1], 3) for k, v in tm.items()}, "launches", b.last_launches)
tot = sum(v[0] / v[1] for v in tm.values())
print("ms/step %.3f -> %.1f Msps (%.1f real-time 20 Msps streams)" % (tot, S * n / tot / 1e3, S * n / tot / 1e3 / 20))
