#!/usr/bin/env python3
"""dsd_in chain at 1024 x 2.4 Msps (development probe; run under `ncu --metrics gpu__time_duration.sum` for a launch list)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdr_pmr446_b200 import chain, synth

S, fs, n = int(os.environ.get("DSD_STREAMS", "1024")), 2400000, int(os.environ.get("DSD_CHUNK", "2400000"))
base = torch.from_numpy(synth.cfg2_capture(n=n)).cuda()
iq = base.unsqueeze(0).repeat(S, 1).contiguous()
d = chain.DsdBatch(n_streams=S, fs_in=fs, in_fmt=1, max_chunk=n)
pcm = torch.empty((S, d.max_out), dtype=torch.int16, device="cuda")
for _ in range(3):
    d.execute_device(iq, n, pcm=pcm)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    d.execute_device(iq, n, pcm=pcm)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("dsd: %.3f ms per step, %.1f Msps" % (ms, S * n / ms / 1e3))
