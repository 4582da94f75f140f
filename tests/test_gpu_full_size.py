"""BASELINE.json's full batch size (1024 streams x 2.4 Msps x 1 s, device-resident like bench.py): properties that do
not need the oracle at that size, plus the oracle on a handful of the streams.

 - chunk invariance: one 2.4 M-sample call and four 600 k-sample calls give the same s16 for every stream up to float32
   rounding of the DC-blocker carry (its segment grid starts at the chunk start): at most 1 LSB on under 1 % of the
   samples of the signal-bearing channels;
 - batch-position invariance: streams fed the same samples give bit-identical rows wherever they sit in the batch;
 - the first eight (distinct) streams equal the CPU oracle within 1 LSB on their signal-bearing channels.
"""
import numpy as np
import pytest

from util import PCM_TOL_LSB, active_channels

pytestmark = pytest.mark.gpu


def test_full_batch_properties():
    import torch
    from oracle import oracle as orc
    from sdr_pmr446_b200 import chain, synth
    S, fs, n = 1024, 2400000, 2400000
    base = [synth.make_cu8(synth.CaptureSpec(fs=float(fs), carriers=synth.rotated_carriers(s)), n, 500 + s) for s in range(8)]
    base_d = torch.stack([torch.from_numpy(b) for b in base]).cuda()
    iq = base_d.repeat(S // 8, 1).contiguous()            # stream s carries capture s % 8
    assert iq.shape == (S, 2 * n)
    b = chain.PmrBatch(n_streams=S, fs_in=fs, in_fmt=1, audio_gain=1.0, max_chunk=n)
    ld = b.max_ns
    one = torch.zeros((S, 16, ld), dtype=torch.int16, device="cuda")
    ny, ns = b.execute_device(iq, n, {"pcm": one, "ld": ld})
    torch.cuda.synchronize()
    assert (ny, ns) == (200000, 12500)
    # four calls, outputs appended
    b.reset()
    four = torch.zeros((S, 16, ld), dtype=torch.int16, device="cuda")
    part = torch.zeros((S, 16, ld), dtype=torch.int16, device="cuda")
    col = 0
    for k in range(4):
        sl = iq[:, 2 * 600000 * k:2 * 600000 * (k + 1)].contiguous()
        _, ns_k = b.execute_device(sl, 600000, {"pcm": part, "ld": ld})
        torch.cuda.synchronize()
        four[:, :, col:col + ns_k] = part[:, :, :ns_k]
        col += ns_k
    b.close()
    assert col == 12500
    # compared on the signal-bearing channels: on a noise-only channel arg() amplifies a rounding difference without bound
    act = torch.zeros((S, 16), dtype=torch.bool)
    for s in range(S):
        act[s, active_channels(synth.rotated_carriers(s % 8))] = True
    act = act.cuda()
    diff = (one[:, :, :12500].to(torch.int32) - four[:, :, :12500].to(torch.int32)).abs()
    diff = diff[act]
    assert int(diff.max()) <= PCM_TOL_LSB
    assert float((diff != 0).float().mean()) < 1e-2
    # same samples, different place in the batch
    ref8 = one[:8, :, :12500]
    for blk in (1, 37, 127):
        assert torch.equal(one[8 * blk:8 * blk + 8, :, :12500], ref8)
    # the distinct streams against the oracle
    g = ref8.cpu().numpy()
    for s in range(8):
        o = orc.PmrOracle(fs_in=fs, in_fmt=1, audio_gain=1.0, chunk=n)
        r = o.execute(base[s], want=("pcm",))
        o.close()
        for c in active_channels(synth.rotated_carriers(s)):
            d = np.abs(g[s, c, 600:12500].astype(np.int32) - r["pcm"][c, 600:12500].astype(np.int32))
            assert d.max() <= PCM_TOL_LSB, (s, c, int(d.max()))
