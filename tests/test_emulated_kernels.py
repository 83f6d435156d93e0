"""The CUDA kernel SOURCES compiled for the host on the SIMT emulator (tests/simt_emu, a test tool:
one CPU thread per CUDA thread) against the oracle: catches kernel-logic regressions in the
CPU-only suite.  Small cases only; the real parity runs are the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, _ptr, judge_float


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "simt_emu")], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    return C.CDLL(os.path.join(ROOT, "tests", "simt_emu", "_build", "libns_emu.so"))


@pytest.fixture(scope="module")
def oracle():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(ROOT, "oracle", "liboracle_ns.so"))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("fs,mode,frames,fpl", [(16000, 2, 70, 16), (8000, 1, 70, 1)])
def test_nsx_kernel_source_bit_exact(emu, oracle, nslib_host_synth, fs, mode, frames, fpl):
    n, fl = 4, fs // 100
    x = nslib_host_synth(n, fs, frames * fl, first_stream=2)   # tone bursts, chirp, silence, clipping
    out = np.zeros_like(x)
    assert emu.emu_nsx_run(fs, mode, 1, n, frames, fpl, _ptr(x), _ptr(out)) == 0
    for s in range(n):
        ref = np.zeros(frames * fl, np.int16)
        xs = np.ascontiguousarray(x[s])
        oracle.nsx_oracle_run(fs, mode, frames, _ptr(xs), _ptr(ref))
        assert np.array_equal(ref, out[s]), "stream %d" % s


@pytest.mark.timeout(300)
@pytest.mark.parametrize("fs,mode,frames,fpl", [(16000, 2, 70, 7), (8000, 3, 70, 70)])
def test_nsf_kernel_source_within_tolerance(emu, oracle, nslib_host_synth, fs, mode, frames, fpl):
    n, fl = 4, fs // 100
    x = nslib_host_synth(n, fs, frames * fl, first_stream=2)
    xf = x.astype(np.float32)
    out = np.zeros_like(xf)
    assert emu.emu_nsf_run(fs, mode, 1, 0, n, frames, fpl, _ptr(xf), _ptr(out), None) == 0
    strict = 0
    for s in range(n):
        ref = np.zeros(frames * fl, np.float32)
        xs = np.ascontiguousarray(x[s])
        oracle.nsf_oracle_run(fs, mode, frames, _ptr(xs), _ptr(ref), None)
        r = judge_float(ref, out[s])
        assert r[1], "stream %d outside envelope: %.3f LSB %.1f dB" % (s, r[2], r[3])
        strict += r[0]
    assert strict >= 3
