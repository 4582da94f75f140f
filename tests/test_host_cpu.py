"""CPU-only checks of the host layer: the liquid-signature header is fully exported by the CUDA library, the shim
refuses to create objects without a GPU, and the C harness built against the CPU oracle reproduces the Python
oracle driver (so the harness source itself is a faithful transcription of the reference loop)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import liquid_api
from oracle import oracle as orc
from sdr_pmr446_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_header_symbols_are_exported():
    txt = open(os.path.join(ROOT, "include", "pmr446_liquid_shim.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    declared = sorted(set(re.findall(r"\b([a-z0-9_]+_(?:crcf|rrrf|s|create|destroy|push|read|write|size|release|execute|reset|set_scale|max_size)\w*)\s*\(", txt)))
    declared = [d for d in declared if not d.endswith("_s")]
    L = C.CDLL(_lib.LIB_PATH)
    assert len(declared) >= 50, declared
    for name in declared:
        assert hasattr(L, name), name
    for name in liquid_api.SHIM_SYMBOLS:
        assert hasattr(L, name), name
        assert name in txt, name


def test_shim_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = liquid_api.bind(C.CDLL(_lib.LIB_PATH))
    assert not L.iirfilt_crcf_create_dc_blocker(0.0005)
    assert not L.msresamp_crcf_create(0.1953125, 60.0)
    assert not L.firpfbch_crcf_create_kaiser(0, 16, 13, 80.0)
    assert not L.freqdem_create(0.5)
    assert not L.asgramcf_create(120)
    # host containers (no arithmetic) still work
    q = L.cbuffercf_create(64)
    v = np.arange(20, dtype=np.complex64)
    assert L.cbuffercf_write(q, v.ctypes.data, 20) == 0 and L.cbuffercf_size(q) == 20
    L.cbuffercf_release(q, 16)
    assert L.cbuffercf_size(q) == 4
    L.cbuffercf_destroy(q)


def test_c_harness_on_oracle_matches_python_driver(tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s", "pmr446_liquid_loop_cpu"])
    n = 250000
    iq8 = synth.make_cu8(synth.CaptureSpec(fs=1024000.0), n, 446)
    x = (iq8.astype(np.float32) - np.float32(127.4)) * np.float32(1.0 / 128.0)
    cf = (x[0::2] + 1j * x[1::2]).astype(np.complex64)
    cap = tmp_path / "cap.cf32"
    cf.tofile(cap)
    out = tmp_path / "a.s16"
    subprocess.check_call([os.path.join(ROOT, "host", "pmr446_liquid_loop_cpu"), "-i", str(cap), "-o", str(out), "-c", "2"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    got = np.fromfile(out, np.int16)
    o = orc.PmrOracle(in_fmt=1, audio_gain=1.0)
    r = o.run(iq8)
    o.close()
    assert got.size == r["ns"]
    assert np.array_equal(got, r["pcm"][1])


def test_python_transcription_of_loop_equals_chain_driver():
    """liquid_api.reference_loop (used for the GPU shim test) == oracle/chains.c on the same input."""
    O = liquid_api.bind(orc.lib())
    hp, lp = liquid_api.reference_taps()
    n = 120000
    iq8 = synth.make_cu8(synth.CaptureSpec(fs=1024000.0), n, 446)
    x = (iq8.astype(np.float32) - np.float32(127.4)) * np.float32(1.0 / 128.0)
    cf = (x[0::2] + 1j * x[1::2]).astype(np.complex64)
    a = liquid_api.reference_loop(O, cf, hp, lp, active_chan=6)
    o = orc.PmrOracle(in_fmt=1, audio_gain=1.0)
    r = o.run(iq8)
    o.close()
    assert np.array_equal(a["res"], r["res"]) and np.array_equal(a["chan"], r["chan"])
    assert np.array_equal(a["audio"], r["audio"][6])


def test_public_headers_are_plain_c_and_cxx():
    """The drop-in boundary is a C ABI: both public headers must compile stand-alone as C99 and as C++."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    for hdr in ("pmr446_b200.h", "pmr446_liquid_shim.h"):
        for lang, std in (("c", "-std=c99"), ("c++", "-std=c++11")):
            subprocess.check_call(["gcc", "-x", lang, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only", os.path.join(inc, hdr)])
