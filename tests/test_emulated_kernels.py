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
def test_nsf_kernel_source_keeps_the_references_state(emu, reflib, nslib_host_synth, fs, mode, frames, fpl):
    """The float kernel source on the emulator against the compiled reference: the whole decision-directed
    state (trackers, densities, noise, gains, LRT, pause spectrum, features, model parameters, prior) is the
    reference struct bit for bit after 70 frames; the output differs only by the inverse FFT's rounding."""
    from conftest import bits_equal, ref_ns_trace
    n, fl = 4, fs // 100
    nb = 129 if fs == 16000 else 65
    x = nslib_host_synth(n, fs, frames * fl, first_stream=2)
    xf = x.astype(np.float32)
    out = np.zeros_like(xf)
    words = emu.emu_nsf_state_words()
    st = np.zeros((n, words), np.uint32)
    assert emu.emu_nsf_run_state(fs, mode, 1, 0, n, frames, fpl, _ptr(xf), _ptr(out), None, _ptr(st)) == 0
    for s in range(n):
        ref, tr = ref_ns_trace(reflib, fs, mode, x[s])
        r = judge_float(ref, out[s])
        assert r[0] and r[2] <= 0.25 and r[3] >= 120.0, "stream %d: %.3f LSB %.1f dB" % (s, r[2], r[3])
        f = st[s].view(np.float32)
        rec = f[548:548 + nb * 12].reshape(nb, 12)
        got = {"lquantile": rec[:, 0:3].T, "density": rec[:, 3:6].T, "quantile": rec[:, 6], "smooth": rec[:, 7],
               "noisePrev": rec[:, 8], "magn": rec[:, 9], "logLrt": rec[:, 10], "magnAvgPause": rec[:, 11],
               "featureData": f[16:23], "priorModelPars": f[8:15]}
        for key, val in got.items():
            assert bits_equal(val, tr[key][-1]), "stream %d: %s is not the reference's" % (s, key)
        assert f[15] == tr["prior"][-1]


def split_signals(synth, n, fs, frames, nb=1):
    """Process signal = stream s; Analyze signal = the same plus an 'echo' (another stream at half
    level) that a canceller would have removed in between.  Frames 20-24 of the Analyze signal and
    40-44 of the Process signal are digital silence (the two zero-energy early-outs, apart)."""
    fl = 80 if fs == 8000 else 160
    a = synth(n, fs, frames * fl, first_stream=2).astype(np.float32)
    e = synth(n, fs, frames * fl, first_stream=11).astype(np.float32)
    x = np.zeros((n, frames, nb, fl), np.float32)
    x[:, :, 0, :] = a.reshape(n, frames, fl)
    for b in range(1, nb):
        x[:, :, b, :] = 0.25 * synth(n, fs, frames * fl, first_stream=20 + b).astype(np.float32).reshape(n, frames, fl)
    ana = (a + 0.5 * e).reshape(n, frames, fl)
    ana[:, 20:25] = 0.0
    x[:, 40:45] = 0.0
    return np.ascontiguousarray(ana.reshape(n, frames * fl)), np.ascontiguousarray(x)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("fs,mode,nb,fpl,fused", [(16000, 2, 1, 9, 0), (16000, 1, 2, 80, 30), (8000, 3, 1, 1, 0)])
def test_nsf_split_kernel_source_against_reference(emu, reflib, nslib_host_synth, fs, mode, nb, fpl, fused):
    """Analyze and Process fed different signals against the compiled
    reference driven through its own API, including a stream handed over from the fused kernel."""
    n, frames = 4, 80
    ana, x = split_signals(nslib_host_synth, n, fs, frames, nb)
    out = np.zeros_like(x)
    assert emu.emu_nsf_run_split(fs, mode, nb, n, frames, fpl, fused, _ptr(ana), _ptr(x), _ptr(out)) == 0
    strict = 0
    for s in range(n):
        ref = reflib.ns_split(fs, mode, ana[s], x[s], nb, fused)
        r = judge_float(ref.ravel(), out[s].ravel())
        assert r[0] and r[2] <= 0.25 and r[3] >= 120.0, "stream %d: %.3f LSB %.1f dB" % (s, r[2], r[3])


@pytest.mark.timeout(300)
@pytest.mark.parametrize("fs,nb", [(16000, 1), (16000, 2)])
def test_nsf_split_kernel_halves_run_apart(emu, reflib, nslib_host_synth, fs, nb):
    """The split kernel issued as an Analyze-half launch and a Process-half launch per frame (the single-stream
    WebRtcNs_Analyze / WebRtcNs_Process pair) gives the output of the one-launch form bit for bit, and the prior
    speech probability sitting in the state between the two halves is the one the reference returns there."""
    n, frames = 3, 60
    ana, x = split_signals(nslib_host_synth, n, fs, frames, nb)
    whole = np.zeros_like(x)
    apart = np.zeros_like(x)
    mid = np.zeros((n, frames), np.float32)
    assert emu.emu_nsf_run_split(fs, 2, nb, n, frames, 1, 0, _ptr(ana), _ptr(x), _ptr(whole)) == 0
    assert emu.emu_nsf_run_split_phased(fs, 2, nb, n, frames, 1, 0, _ptr(ana), _ptr(x), _ptr(apart), 1, _ptr(mid)) == 0
    assert np.array_equal(whole.view(np.uint32), apart.view(np.uint32))
    if nb == 1:
        # same signal on both sides: the reference's getter between Analyze and Process
        x1 = nslib_host_synth(n, fs, frames * 160, first_stream=4)
        xf = x1.astype(np.float32).reshape(n, frames, 1, 160)
        out = np.zeros_like(xf)
        assert emu.emu_nsf_run_split_phased(fs, 2, 1, n, frames, 1, 0, _ptr(np.ascontiguousarray(xf.reshape(n, -1))), _ptr(xf), _ptr(out), 1, _ptr(mid)) == 0
        for s in range(n):
            want_mid = np.zeros(frames, np.float32)
            want_out = np.zeros(frames * 160, np.float32)
            assert reflib.lib.ref_ns_prior_between(fs, 2, frames, _ptr(np.ascontiguousarray(x1[s])), _ptr(want_out), _ptr(want_mid)) == 0
            assert np.array_equal(mid[s].view(np.uint32), want_mid.view(np.uint32))
            assert np.abs(out[s].ravel() - want_out).max() <= 0.25
