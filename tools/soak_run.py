#!/usr/bin/env python3
"""Long-run index test (run by hand on a GPU box): one 2.4 Msps cu8 stream pushed past 2^31 input samples by feeding
the same 2.4 M-sample chunk over and over; the per-chunk counts must equal the CPU oracle's for every chunk and the
s16 audio of the last chunks must match within 1 LSB (absolute indices, 32-bit relative arithmetic, ring wrap-around)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle as orc
from sdr_pmr446_b200 import chain, synth

fs, n = 2400000, 2400000
chunks = int(os.environ.get("SOAK_CHUNKS", "900"))          # 900 x 2.4 M = 2.16e9 > 2^31
car = synth.rotated_carriers(1)
iq = synth.make_cu8(synth.CaptureSpec(fs=float(fs), carriers=car), n, 446)
gpu = chain.PmrBatch(n_streams=1, fs_in=fs, in_fmt=1, audio_gain=1.0, max_chunk=n)
ref = orc.PmrOracle(fs_in=fs, in_fmt=1, audio_gain=1.0, chunk=n)
act = sorted({c.channel - 1 for c in car})
t0 = time.time()
worst = 0
for k in range(chunks):
    g = gpu.execute(iq[None, :], want=("pcm",))
    r = ref.execute(iq, want=("pcm",))
    assert g["ny"] == r["ny"] and g["ns"] == r["ns"], (k, g["ny"], r["ny"], g["ns"], r["ns"])
    if k % 100 == 99 or k >= chunks - 3:
        for c in act:
            d = np.abs(g["pcm"][0, c].astype(np.int32) - r["pcm"][c].astype(np.int32)).max()
            worst = max(worst, int(d))
            assert d <= 1, (k, c, int(d))
        print("chunk %d (%.2e input samples): counts equal, worst pcm difference %d LSB, %.0f s" % (k + 1, (k + 1) * float(n), worst, time.time() - t0),
              flush=True)
print("soak ok: %d chunks" % chunks)
