#!/usr/bin/env python3
"""Development probe: device-resident timing of the batched PMR chain, per kernel group (not the judged bench)."""
import argparse
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sdr_pmr446_b200 import chain, synth

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=256)
ap.add_argument("--fs", type=int, default=2400000)
ap.add_argument("--chunk", type=int, default=2400000)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--want", default="pcm")
a = ap.parse_args()

S, n = a.streams, a.chunk
base = [synth.make_cu8(synth.CaptureSpec(fs=float(a.fs), carriers=synth.rotated_carriers(s)), n, 446 + s) for s in range(4)]
iq = torch.empty((S, 2 * n), dtype=torch.uint8, device="cuda")
for s in range(S):
    iq[s] = torch.from_numpy(np.roll(base[s % 4], 2 * (s // 4) * 16)).cuda()
b = chain.PmrBatch(n_streams=S, fs_in=a.fs, in_fmt=1, audio_gain=1.0, max_chunk=n)
ld = b.max_ns
outs = {"ld": ld}
if "pcm" in a.want:
    outs["pcm"] = torch.empty((S, 16, ld), dtype=torch.int16, device="cuda")
if "audio" in a.want:
    outs["audio"] = torch.empty((S, 16, ld), dtype=torch.float32, device="cuda")
print("fp32 peak TFLOP/s", chain.measure_fp32_peak())
b.timing(True)
for it in range(a.steps + 2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ny, ns = b.execute_device(iq, n, outs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("step %d: %.3f ms  %.1f Msps  ny=%d ns=%d launches=%d" % (it, ms, S * n / ms / 1e3, ny, ns, b.last_launches))
tm = b.get_timings()
print("per-kernel avg ms:", {k: round(v[0] / v[1], 3) for k, v in tm.items()})
