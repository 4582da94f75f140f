#!/usr/bin/env python3
"""Per-kernel timing of the bench workload with the waterfall switched on (development probe)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sdr_pmr446_b200 import chain, synth

S, fs, n, W = int(os.environ.get("WF_STREAMS", "1024")), 2400000, 2400000, int(os.environ.get("WF_W", "120"))
base = torch.from_numpy(synth.make_cu8(synth.CaptureSpec(fs=float(fs)), n, 446)).cuda()
iq = base.unsqueeze(0).repeat(S, 1).contiguous()
b = chain.PmrBatch(n_streams=S, fs_in=fs, in_fmt=1, waterfall=W, audio_gain=1.0, max_chunk=n)
outs = {"pcm": torch.empty((S, 16, b.max_ns), dtype=torch.int16, device="cuda"), "ld": b.max_ns,
        "ascii": torch.empty((S, W), dtype=torch.uint8, device="cuda"), "peak": torch.empty((S, 2), dtype=torch.float32, device="cuda")}
for _ in range(2):
    b.execute_device(iq, n, outs)
b.timing(True)
for _ in range(3):
    b.execute_device(iq, n, outs)
tm = b.get_timings()
print({k: round(v[0] / v[1], 3) for k, v in tm.items()})
