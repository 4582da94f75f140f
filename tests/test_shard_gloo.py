"""world_size-2 test of the multi-GPU path on CPU (gloo): streams are sharded, every rank runs the chain on its shard
(here: the CPU oracle stands in for the per-rank device work), statistics are all-gathered, and the sharded result
equals the single-process result stream by stream."""
import os
import socket

import numpy as np
import pytest

from sdr_pmr446_b200 import shard, synth

FS, N, STREAMS = 1024000, 60000, 5


def _stream_pcm(s):
    from oracle import oracle as orc
    iq = synth.make_cu8(synth.CaptureSpec(fs=float(FS), carriers=synth.rotated_carriers(s)), N, 446 + s)
    o = orc.PmrOracle(fs_in=FS, in_fmt=1, audio_gain=1.0, chunk=N)
    r = o.run(iq, N)
    o.close()
    return r["pcm"]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = shard.shard_range(STREAMS, world, rank)
    sums = [shard.stream_checksum(_stream_pcm(s)) for s in range(start, start + count)]
    stats = shard.gather_stats({"streams": count, "samples": count * N, "checksum": float(sum(c % (1 << 40) for c in sums)),
                                "elapsed_ms": 10.0 * (rank + 1)})
    worst = shard.max_over_ranks(10.0 * (rank + 1))
    q.put((rank, start, count, sums, stats, worst))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_all_streams():
    for total in (1, 5, 8, 1024, 8192):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                s, c = shard.shard_range(total, world, r)
                seen += list(range(s, s + c))
            assert seen == list(range(total))
    assert shard.shard_range(8192, 8, 3) == (3072, 1024)


def test_two_rank_shard_equals_single_process():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = [shard.stream_checksum(_stream_pcm(s)) for s in range(STREAMS)]
    got = []
    for rank, start, count, sums, stats, worst in results:
        got += sums
        assert worst == 20.0                                  # max over ranks
        assert [int(st["streams"]) for st in stats] == [3, 2]  # every rank sees every rank's statistics
        assert sum(int(st["samples"]) for st in stats) == STREAMS * N
    assert got == single
