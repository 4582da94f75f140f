"""On-hardware shard equivalence: the streams of a batch are split over two GPUs, one process
per GPU with torch.distributed over NCCL, every rank runs the CUDA chain (C ABI) on its shard, the s16 audio is
all-gathered over NCCL, and each stream's result must be BITWISE equal to the single-GPU run of the whole batch.
Needs two GPUs (`gpurun --gpus 2`); skipped otherwise.  tests/test_shard_gloo.py covers the same host logic on CPU."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FS, N, CHUNK, STREAMS = 2400000, 480000, 240000, 6


def _captures(lo, hi):
    from sdr_pmr446_b200 import synth
    return np.stack([synth.make_cu8(synth.CaptureSpec(fs=float(FS), carriers=synth.rotated_carriers(s)), N, 446 + s) for s in range(lo, hi)])


def _run(iq, device):
    from sdr_pmr446_b200 import chain
    b = chain.PmrBatch(n_streams=iq.shape[0], device=device, fs_in=FS, in_fmt=1, audio_gain=1.0, max_chunk=CHUNK)
    g = b.run(iq, CHUNK, want=("pcm", "demod"))
    b.close()
    return g


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:   # the parent must not sit out its queue timeout on a GPU box
        import traceback
        q.put(("error", "rank %d: %s\n%s" % (rank, e, traceback.format_exc())))


def _worker_body(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from sdr_pmr446_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    start, count = shard.shard_range(STREAMS, world, rank)
    g = _run(_captures(start, start + count), rank)
    ns = g["ns"]
    # NCCL all-gather of the shard's audio (padded to the largest shard) and of its discriminator output
    most = max(shard.shard_range(STREAMS, world, r)[1] for r in range(world))
    pcm = torch.zeros((most, 16, ns), dtype=torch.int16, device="cuda")
    pcm[:count] = torch.from_numpy(g["pcm"]).cuda()
    dem = torch.zeros((most, 16, ns), dtype=torch.float32, device="cuda")
    dem[:count] = torch.from_numpy(g["demod"]).cuda()
    pcm_b = pcm.view(torch.uint8)                     # NCCL has no int16 type: the s16 audio travels as bytes
    all_pcm_b = [torch.empty_like(pcm_b) for _ in range(world)]
    all_dem = [torch.empty_like(dem) for _ in range(world)]
    dist.all_gather(all_pcm_b, pcm_b)
    dist.all_gather(all_dem, dem)
    all_pcm = [t.view(torch.int16) for t in all_pcm_b]
    stats = shard.gather_stats({"streams": count, "samples": count * N}, device="cuda")
    if rank == 0:
        rows_p, rows_d = [], []
        for r in range(world):
            c = shard.shard_range(STREAMS, world, r)[1]
            rows_p.append(all_pcm[r][:c].cpu().numpy())
            rows_d.append(all_dem[r][:c].cpu().numpy())
        q.put((np.concatenate(rows_p), np.concatenate(rows_d), ns, stats))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shard_bitwise_equals_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    if isinstance(got[0], str):
        for p in procs:
            p.kill()
        pytest.fail(got[1])
    pcm2, dem2, ns2, stats = got
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = _run(_captures(0, STREAMS), 0)
    assert single["ns"] == ns2
    assert [int(st["streams"]) for st in stats] == [3, 3]
    assert np.array_equal(single["pcm"], pcm2)                                       # bitwise, every stream and channel
    assert np.array_equal(single["demod"].view(np.uint32), dem2.view(np.uint32))     # float outputs bit for bit too
