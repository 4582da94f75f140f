#!/usr/bin/env python3
"""Benchmark of the PMR446 receive chain on B200 (BASELINE.json metric: IQ Msamples/s through the
full chain + roofline fraction).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[2]; configs[4] is the same shard per GPU at N = 8): a batch of
STREAMS = 1024 independent 2.4 Msps cu8 PMR446 captures per GPU, all 16 channels demodulated to s16
audio.  One step = one second of signal of every stream (2.4 M samples x 1024 streams = 2.46 G
samples, 4.9 GB of input per GPU -- far larger than the 126 MB L2, so no L2 flush is needed between
steps).  Filter state carries from step to step exactly as in streaming use.

  value     device-resident throughput: inputs already in HBM, s16 audio left in HBM; K steps timed with
            CUDA events on the launch stream between barriers, max over ranks.
  e2e       the same K steps through the host-buffer C-ABI call pmr446_batch_execute(): pinned host IQ
            -> H2D -> chain -> D2H of the s16 audio, copies inside the timed region.
  roofline  the kernel with the longest launch, timed live with CUDA events inside the timed steps: algorithmic
            bytes / measured HBM peak and algorithmic flop / measured FP32 FFMA peak, reported for the roofline that
            binds (the larger fraction); "kernels" carries the same numbers for every launch of the step.
  cpu_baseline  the CPU oracle (port of the reference chain, all 16 channels) on all host cores.

--impl reference times the reference's CPU chain (the oracle port -- the reference itself cannot be
built here) on the host cores with the same config/metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FS = 2400000
CHUNK = 2400000            # one second of signal per step
STREAMS = int(os.environ.get("PMR446_BENCH_STREAMS", "1024"))
METRIC = "IQ Msamples/s through full PMR446 chain"
UNIT = "Msamples/s"
# algorithmic work of the 2.4 Msps / 16-channel chain
FLOP_PER_SAMPLE = 122.1
BYTES_PER_SAMPLE = 2.17
# per kernel launch: algorithmic flop and HBM bytes per INPUT sample (SURVEY.md 8d table, 2.4 Msps column, divided
# by 2.4e6; direct-form FIR counts; bytes = what the launch must read + write once), and what it replaces
KERNELS = {
    "cascade0": {"what": "cascade_kernel<cu8, DC, m=3, m=5>: cu8 load, DC blocker, two half-band decimators -> 600 kHz ring",
                 "flop": (19.2 + 31.2 + 25.2) / 2.4, "bytes": 2.0 + 0.6e6 * 8 / 2.4e6},
    "cascade1": {"what": "hbarb_tile_kernel<10, 2, 3>: m=10 half-band decimator + 14-tap arbitrary resampler -> 200 kHz ring",
                 "flop": (24.6 + 11.2) / 2.4, "bytes": 0.6e6 * 8 / 2.4e6 + 0.2e6 * 8 / 2.4e6},
    "channelize": {"what": "channelize16_kernel: NCO mix, 16 x 26-tap polyphase bank, 16-point DFT, FM discriminator",
                   "flop": (1.2 + 20.8 + 4.0 + 3.8) / 2.4, "bytes": 0.2e6 * 8 / 2.4e6 + 0.2e6 * 4 / 2.4e6},
    "audio": {"what": "audio_fft_kernel: 377-tap CTCSS high-pass FIR + gain + de-emphasis + s16 by 4096-point overlap-save "
                      "(flop = direct-form count of SURVEY 8d; the kernel executes ~10x fewer)",
              "flop": (150.8 + 1.0) / 2.4, "bytes": 0.2e6 * 4 / 2.4e6 + 0.2e6 * 2 / 2.4e6},
}
WORKLOAD = "configs[2]: %d independent 2.4 Msps cu8 PMR446 captures x 16 channels per GPU, 1 s of signal per step" % STREAMS


emit = None   # set by main(): prints the one JSON line on the real stdout


def config_dict(n_gpus):
    return {"workload": WORKLOAD, "streams_per_gpu": STREAMS, "fs_in": FS, "samples_per_stream_per_step": CHUNK,
            "channels": 16, "in_fmt": "cu8", "out": "s16 audio, 16 x 12.5 kHz per stream", "parallelism": "streams sharded, %d GPU(s)" % n_gpus,
            "l2": "inputs (4.9 GB/step/GPU) larger than L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
def make_base_captures(count, n):
    from sdr_pmr446_b200 import synth
    return [synth.make_cu8(synth.CaptureSpec(fs=float(FS), carriers=synth.rotated_carriers(s)), n, 446 + s) for s in range(count)]


def cpu_chain_rate(threads, seconds_of_signal, active_only=-1):
    """Oracle (CPU port of the reference chain) on `threads` host threads, one stream each.  Returns
    (Msamples/s aggregate, wall seconds)."""
    from oracle import oracle as orc
    base = make_base_captures(min(threads, 4), CHUNK)
    objs = [orc.PmrOracle(fs_in=FS, in_fmt=1, audio_gain=1.0, chunk=CHUNK, active_only=active_only) for _ in range(threads)]

    def work(i):
        for _ in range(seconds_of_signal):
            objs[i].execute(base[i % len(base)], want=("pcm",))

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    for o in objs:
        o.close()
    return threads * seconds_of_signal * CHUNK / dt / 1e6, dt


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], None, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        top = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (top[len(top) // 2] if top else None), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    secs = 4
    vals = []
    for it in range(args.warmup + args.steps):
        v, dt = cpu_chain_rate(cores, secs)
        if it >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    sample = "%d host threads x %d s of one 2.4 Msps stream each per step (all 16 channels demodulated)" % (cores, secs)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = CPU oracle port of the liquid-dsp chain (reference cannot be compiled here: liquid-dsp v1.7.0 absent)"}
    emit(line)


def bind_near_gpu(local_rank):
    """Pins this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers of the
    end-to-end arm are allocated next to the GPU's PCIe root (first-touch).  Returns the node or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from sdr_pmr446_b200 import chain

    torch.cuda.set_device(local_rank)
    numa_node = bind_near_gpu(local_rank) if world > 1 else None
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)

    S, n = STREAMS, CHUNK
    # synthetic captures: 8 distinct seeded captures, circularly shifted per stream
    base = make_base_captures(8, n)
    base_t = [torch.from_numpy(b) for b in base]
    iq_host = torch.empty((S, 2 * n), dtype=torch.uint8).pin_memory()
    for s in range(S):
        sh = 2 * 16 * ((s // 8) + rank * (S // 8))
        iq_host[s, :2 * n - sh] = base_t[s % 8][sh:]
        iq_host[s, 2 * n - sh:] = base_t[s % 8][:sh]
    iq_dev = iq_host.to(dev, non_blocking=True)
    batch = chain.PmrBatch(n_streams=S, device=local_rank, fs_in=FS, in_fmt=1, audio_gain=1.0, max_chunk=n)
    ld = batch.max_ns
    pcm_dev = torch.empty((S, 16, ld), dtype=torch.int16, device=dev)
    outs = {"pcm": pcm_dev, "ld": ld}
    torch.cuda.synchronize()
    fp32_peak = chain.measure_fp32_peak()

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm ("value") ----------------
    launches = 0
    for _ in range(args.warmup):
        batch.execute_device(iq_dev, n, outs)
    batch.timing(True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        batch.execute_device(iq_dev, n, outs)
        launches += batch.last_launches
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    timings = batch.get_timings()
    batch.timing(False)
    checksum = int(pcm_dev[:, :, :12500].to(torch.int64).abs().sum().item())

    # ---------------- end-to-end arm ("e2e"): host buffers through the C ABI ----------------
    pcm_host = np.zeros((S, 16, ld), np.int16)
    pcm_host_t = torch.from_numpy(pcm_host).pin_memory()
    pcm_host = pcm_host_t.numpy()
    iq_np = iq_host.numpy()
    import ctypes as C
    from sdr_pmr446_b200._lib import Outputs, check, lib
    o = Outputs()
    o.ld, o.res_ld, o.pcm = ld, batch.max_res, pcm_host.ctypes.data
    ny, ns = C.c_uint(0), C.c_uint(0)
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        check(lib().pmr446_batch_execute(batch.h, iq_np.ctypes.data, iq_np.strides[0], n, C.byref(o), C.byref(ny), C.byref(ns)),
              "pmr446_batch_execute")

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_ns = ns.value
    clocks = sampler.stop()   # sampled across both timed regions

    # ---------------- reduce over ranks ----------------
    # NCCL is used for statistics only; the data path has no collective (streams are independent)
    from sdr_pmr446_b200 import shard
    rank_stats = shard.gather_stats({"samples": float(S) * n * args.steps, "elapsed_ms": ms_total, "e2e_ms": e2e_s * 1e3,
                                     "checksum": float(checksum % (1 << 40)), "sm_mhz": float(clocks.get("sm_mhz") or 0)}, device=dev)
    ms_total = shard.max_over_ranks(ms_total, device=dev)
    e2e_ms = shard.max_over_ranks(e2e_s * 1e3, device=dev)
    total_samples = float(world) * S * n * args.steps
    value = total_samples / (ms_total * 1e-3) / 1e6
    e2e_value = float(world) * S * n * e2e_steps / (e2e_ms * 1e-3) / 1e6

    if rank == 0:
        peaks, peak_src = measured_peaks()
        step_ms = ms_total / args.steps
        shares = {k: round(v[0] / max(v[1], 1) / step_ms, 4) for k, v in timings.items()}
        hbm_peak = peaks.get("hbm_gbs")
        traffic_all = {}
        tpath = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic_all = json.load(f)
        per_kernel = {}
        for name, k in KERNELS.items():
            ms, cnt = timings.get(name, (0.0, 0))
            if not cnt:
                continue
            avg = ms / cnt                                  # one launch = one step's samples of this rank
            tf = k["flop"] * float(S) * n / (avg * 1e-3) / 1e12
            gbs = k["bytes"] * float(S) * n / (avg * 1e-3) / 1e9
            per_kernel[name] = {"avg_launch_ms": avg, "share_of_step": shares.get(name), "tflops": tf, "fp32_frac": tf / fp32_peak,
                                "gbs": gbs, "hbm_frac": gbs / hbm_peak}
        dom = max(per_kernel, key=lambda q: per_kernel[q]["avg_launch_ms"])
        dk = per_kernel[dom]
        dom_hbm = dk["hbm_frac"] >= dk["fp32_frac"]
        tr = traffic_all.get(dom)
        roofline = {"kernel": dom + " -- " + KERNELS[dom]["what"], "bound": "hbm" if dom_hbm else "fp32",
                    "achieved": dk["gbs"] if dom_hbm else dk["tflops"], "peak": hbm_peak if dom_hbm else fp32_peak,
                    "unit": "GB/s" if dom_hbm else "TFLOP/s", "frac": dk["hbm_frac"] if dom_hbm else dk["fp32_frac"],
                    "traffic": tr * (float(S) / 1024.0) if tr else None,
                    "algorithmic_bytes_per_launch": KERNELS[dom]["bytes"] * float(S) * n,
                    "algorithmic_flop_per_launch": KERNELS[dom]["flop"] * float(S) * n,
                    "peak_source": (peak_src if dom_hbm else "FFMA issue peak measured live by pmr446_measure_fp32_peak "
                                    "(MEASURED_PEAKS.json has no FP32 figure)"),
                    "other_roofline_frac": dk["fp32_frac"] if dom_hbm else dk["hbm_frac"],
                    "avg_launch_ms": dk["avg_launch_ms"], "share_of_step": dk["share_of_step"]}
        cores = os.cpu_count() or 1
        cpu_v, cpu_dt = cpu_chain_rate(cores, 8)
        cpu_f, cpu_fdt = cpu_chain_rate(cores, 8, active_only=1)   # what the reference executes: one demodulated channel
        cpu_1, _ = cpu_chain_rate(1, 8)                             # one stream on one core
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": S * n * 2, "d2h_bytes_per_step": S * 16 * int(e2e_ns) * 2,
                    "steps": e2e_steps, "api": "pmr446_batch_execute (host buffers, pinned)",
                    "numa_bound": numa_node is not None},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "kernels": per_kernel,
            "chain_roofline": {"fp32": {"flop_per_sample": FLOP_PER_SAMPLE, "achieved_tflops": value * 1e6 * FLOP_PER_SAMPLE / 1e12 / world,
                                        "peak_tflops": fp32_peak, "frac": value * 1e6 * FLOP_PER_SAMPLE / 1e12 / world / fp32_peak},
                               "hbm": {"bytes_per_sample": BYTES_PER_SAMPLE, "achieved_gbs": value * 1e6 * BYTES_PER_SAMPLE / 1e9 / world,
                                       "peak_gbs": hbm_peak, "frac": value * 1e6 * BYTES_PER_SAMPLE / 1e9 / world / hbm_peak,
                                       "peak_source": peak_src}},
            "kernel_share_of_step": shares,
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d host threads x 8 s of one 2.4 Msps stream each (all 16 channels demodulated), %.1f s wall" % (cores, cpu_dt),
                             "single_core_value": cpu_1,
                             "reference_faithful_value": cpu_f,
                             "reference_faithful_sample": "same, only one channel demodulated as the reference does (SURVEY 8d), %.1f s wall" % cpu_fdt},
            "checksum": checksum,
            "per_rank": rank_stats,
        }
        emit(line)
    batch.close()
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: anything libraries print meanwhile (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global emit

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
