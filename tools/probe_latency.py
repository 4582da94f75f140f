#!/usr/bin/env python3
"""Single-stream latency of the reference-shaped calls (development probe): one 100 000-sample chunk of 1.024 Msps cu8
through pmr446_receiver_execute / pmr446_batch_execute (host buffers), and the literal liquid-API loop (host/) on the
GPU library vs the CPU oracle."""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from sdr_pmr446_b200 import chain, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fs, chunk, nchunks = 1024000, 100000, 60
iq = synth.cfg1_capture(n=chunk * nchunks, fs=fs)


def timed(obj, fn):
    fn(iq[None, :2 * chunk])
    t = []
    for k in range(1, nchunks):
        t0 = time.perf_counter()
        fn(iq[None, 2 * chunk * k:2 * chunk * (k + 1)])
        t.append(time.perf_counter() - t0)
    obj.close()
    t = np.array(t) * 1e3
    return float(np.median(t)), float(t.max())


rx = chain.PmrReceiver(n_streams=1, fs_in=fs, in_fmt=1, max_chunk=chunk)
print("pmr446_receiver_execute: median %.3f ms, max %.3f ms per 97.7 ms chunk" % timed(rx, rx.execute))
b = chain.PmrBatch(n_streams=1, fs_in=fs, in_fmt=1, max_chunk=chunk)
print("pmr446_batch_execute (16 channels -> s16): median %.3f ms, max %.3f ms" % timed(b, lambda x: b.execute(x, want=("pcm",))))
subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])
cap = "/tmp/lat_cap.cf32"
synth.make_cf32(synth.CaptureSpec(fs=float(fs)), chunk * 30, 446).tofile(cap)
for kind in ("gpu", "cpu"):
    t0 = time.perf_counter()
    subprocess.check_call([os.path.join(ROOT, "host", "pmr446_liquid_loop_" + kind), "-i", cap, "-o", "/tmp/lat_out.s16", "-c", "8"],
                          stdout=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    print("liquid-API loop (%s): %.2f s for %.2f s of signal" % (kind, dt, 30 * chunk / fs))
