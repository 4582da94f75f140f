#!/usr/bin/env python3
"""Development probe: one receiver-mode step (1024 x 2.4 Msps) -- run under `ncu --metrics gpu__time_duration.sum` for the launch list."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sdr_pmr446_b200 import _lib, chain, synth

S, fs, n = int(os.environ.get("RX_STREAMS", "1024")), 2400000, 2400000
base = torch.from_numpy(synth.make_cu8(synth.CaptureSpec(fs=float(fs)), n, 446)).cuda()
iq = base.unsqueeze(0).repeat(S, 1).contiguous()
rx = chain.PmrReceiver(n_streams=S, fs_in=fs, in_fmt=1, max_chunk=n, audio_gain=1.0)
o = {"ld": rx.max_ns, "pcm": torch.empty((S, rx.max_ns), dtype=torch.int16, device="cuda"),
     "status": torch.empty((S, C.sizeof(_lib.RxStatus)), dtype=torch.uint8, device="cuda"),
     "rssi": torch.empty((S, 16), dtype=torch.float32, device="cuda")}
for _ in range(3):
    rx.execute_device(iq, n, o)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    rx.execute_device(iq, n, o)
e1.record()
torch.cuda.synchronize()
print("receiver: %.3f ms per step, %d launches" % (e0.elapsed_time(e1) / 3, rx.last_launches))
