import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box)")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RefLib:
    """oracle/_ref/libns_ref.so: the unmodified reference compiled by oracle/Makefile."""

    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.lib.ref_run_mt.restype = C.c_double

    def ns(self, fs, mode, pcm):
        """pcm int16 [samples] -> (float out or None, int16 out, prior prob per frame)"""
        fl = fs // 100
        nfr = len(pcm) // fl
        x = np.ascontiguousarray(pcm, np.int16)
        of = np.zeros(nfr * fl, np.float32)
        oi = np.zeros(nfr * fl, np.int16)
        pp = np.zeros(nfr, np.float32)
        rc = self.lib.ref_ns_run(fs, mode, nfr, _ptr(x), _ptr(of) if fs <= 16000 else None, _ptr(oi), _ptr(pp))
        assert rc == 0
        return (of if fs <= 16000 else None), oi, pp

    def ns_split(self, fs, mode, ana, x, nb=1, fused_frames=0):
        """Analyze fed `ana` [frames*fl] float, Process fed `x` [frames][nb][fl] float -> float out like x."""
        fl = 80 if fs == 8000 else 160
        ana = np.ascontiguousarray(ana, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        nfr = ana.size // fl
        out = np.zeros_like(x)
        assert self.lib.ref_ns_split_run(fs, mode, nb, nfr, fused_frames, _ptr(ana), _ptr(x), _ptr(out)) == 0
        return out

    def nsx(self, fs, mode, pcm):
        fl = fs // 100
        nfr = len(pcm) // fl
        x = np.ascontiguousarray(pcm, np.int16)
        oi = np.zeros(nfr * fl, np.int16)
        assert self.lib.ref_nsx_run(fs, mode, nfr, _ptr(x), _ptr(oi)) == 0
        return oi


@pytest.fixture(scope="session")
def reflib():
    path = os.path.join(ROOT, "oracle", "_ref", "libns_ref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref", "-j8"],
                                  stdout=subprocess.DEVNULL)
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return RefLib(path)


@pytest.fixture(scope="session")
def reflib_fma():
    """The same reference with FMA contraction in ns_core.c/fft4g.c (sensitivity probe)."""
    path = os.path.join(ROOT, "oracle", "_ref", "libns_ref_fma.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libns_ref_fma.so not built")
    return RefLib(path)


@pytest.fixture(scope="session")
def nslib_host_synth():
    """Host-side synthetic PCM generator of the product library (no GPU needed)."""
    import audiosignalprocess_b200 as pkg
    return pkg.synth_pcm_host


@pytest.fixture(scope="session")
def nslib():
    import audiosignalprocess_b200 as pkg
    return pkg


def snr_db(ref, out):
    ref = np.asarray(ref, np.float64)
    d = np.asarray(out, np.float64) - ref
    den = float((d * d).sum())
    num = float((ref * ref).sum())
    if den == 0.0:
        return 300.0
    if num == 0.0:
        return -300.0
    return 10.0 * np.log10(num / den)


# ---- float parity criteria ---------------------------------------------------------------
# STRICT is BASELINE.json's tolerance: per stream max abs <= 1e-4 full scale and SNR of the difference
# >= 90 dB.  Every GPU float test requires it of EVERY stream (min_strict_frac = 1.0), and the tests of the
# float-sample interfaces additionally hold the kernel to what it actually delivers (FLOAT_MAX_ABS /
# FLOAT_MIN_SNR below; measured on the B200 over 256 streams x 60 s at 8 and 16 kHz: worst 0.027 LSB,
# 133.9 dB -- profiles/r2_float_parity.md).
# Why that is reachable: the float algorithm branches on comparisons decided by the last bit of its
# operands (|lmagn - lquantile| < WIDTH, lmagn > lquantile, speechProb > PROB_RANGE, histogram bins), so
# two conforming builds of the REFERENCE ITSELF (plain vs FMA-contracted, oracle/Makefile) leave each
# other by hundreds of LSB on most 60 s streams (test_reference_sensitivity.py).  The kernel therefore
# keeps the whole decision-directed recursion bit-identical to the plain build (test_gpu_float_exact_state.py);
# only the synthesis (inverse FFT, window, overlap-add) rounds differently.
#   envelope -- SNR >= 55 dB and max abs <= 1e-2 FS: what two builds of the reference show against each
#               other; kept as the outer gate (a failure there is a bug, not a rounding).
STRICT_MAX_ABS = 1e-4 * 32768.0
STRICT_MIN_SNR = 90.0
ENVELOPE_MAX_ABS = 1e-2 * 32768.0
ENVELOPE_MIN_SNR = 55.0
FLOAT_MAX_ABS = 0.25     # LSB, float-sample outputs at 8/16 kHz
FLOAT_MIN_SNR = 120.0    # dB


def judge_float(ref, out, slack=0.0):
    """(strict_ok, envelope_ok, max_abs, snr). slack: extra LSB for int16-rounded outputs."""
    ref = np.asarray(ref, np.float64)
    out = np.asarray(out, np.float64)
    err = float(np.abs(out - ref).max()) if len(ref) else 0.0
    snr = snr_db(ref, out)
    silent = float(np.abs(ref).max()) == 0.0
    strict = err <= STRICT_MAX_ABS + slack and (snr >= STRICT_MIN_SNR - (30.0 if slack else 0.0) or silent)
    env = err <= ENVELOPE_MAX_ABS and (snr >= ENVELOPE_MIN_SNR - (5.0 if slack else 0.0) or silent)
    return strict, env, err, snr


def summarize_parity(results, what, min_strict_frac=1.0, max_abs=None, min_snr=None):
    """results: list of judge_float tuples.  Asserts the gates (strict for at least min_strict_frac of the
    streams -- 1.0 everywhere on the GPU --, envelope for all, and optionally a tighter max abs [LSB] /
    min SNR [dB] for every stream); returns the printable line."""
    n = len(results)
    n_strict = sum(1 for r in results if r[0])
    worst_err = max(r[2] for r in results)
    worst_snr = min(r[3] for r in results)
    line = "%s: %d/%d streams within 1e-4 FS & 90 dB; worst max-abs %.3f LSB, worst SNR %.1f dB" % (
        what, n_strict, n, worst_err, worst_snr)
    print(line)
    assert all(r[1] for r in results), "outside the reference's own cross-build envelope: " + line
    assert n_strict >= min_strict_frac * n, "too few streams within the strict tolerance: " + line
    if max_abs is not None:
        assert worst_err <= max_abs, "max abs above %.3f LSB: %s" % (max_abs, line)
    if min_snr is not None:
        assert worst_snr >= min_snr, "SNR below %.1f dB: %s" % (min_snr, line)
    return line


# ---- decision state of the float suppressor, reference vs GPU -------------------------------------------
# The reference struct after a run (oracle/ref_shim.cc ref_ns_trace) and the GPU state slab of a handle
# (WebRtcNsB200_ExportState; layout csrc/nsf_layout.h behind the 696-byte StateBlobHeader of ns_capi.cu).
NSF_BLOB_HEADER_BYTES = 696
NSF_OFF_BINS, NSF_BIN_REC = 548, 12


def ref_ns_trace(reflib, fs, mode, pcm):
    """-> (float out [samples], dict of per-frame arrays: lquantile/density [frames,3,bins], magn, quantile,
    smooth, noisePrev, logLrt, magnAvgPause [frames,bins], prior [frames])"""
    lib = reflib.lib
    fl = fs // 100
    nfr = len(pcm) // fl
    nb = 129 if fs == 16000 else 65
    W = lib.ref_ns_trace_words()
    tr = np.zeros((nfr, W), np.float32)
    out = np.zeros(nfr * fl, np.float32)
    x = np.ascontiguousarray(pcm, np.int16)
    assert lib.ref_ns_trace(fs, mode, nfr, _ptr(x), _ptr(out), _ptr(tr)) == 0
    per = tr[:, 774:774 + 6 * 129].reshape(nfr, 6, 129)[:, :, :nb]
    u = tr[:, 774 + 6 * 129:]
    return out, {"lquantile": tr[:, 0:387].reshape(nfr, 3, 129)[:, :, :nb], "density": tr[:, 387:774].reshape(nfr, 3, 129)[:, :, :nb],
                 "quantile": per[:, 0], "smooth": per[:, 1], "noisePrev": per[:, 2], "magn": per[:, 3], "logLrt": per[:, 4],
                 "magnAvgPause": per[:, 5], "featureData": u[:, 0:7], "priorModelPars": u[:, 7:14], "prior": u[:, 14]}


def gpu_nsf_state(lib, handle, fs):
    """The per-bin records of one float-NS handle: dict like ref_ns_trace's, one frame."""
    nb = 129 if fs == 16000 else 65
    size = lib.WebRtcNsB200_StateSize(handle)
    buf = np.zeros(size, np.uint8)
    assert lib.WebRtcNsB200_ExportState(handle, _ptr(buf), size) == 0, lib.WebRtcNsB200_LastError()
    st = buf[NSF_BLOB_HEADER_BYTES:].view(np.float32)
    rec = st[NSF_OFF_BINS:NSF_OFF_BINS + nb * NSF_BIN_REC].reshape(nb, NSF_BIN_REC)
    return {"lquantile": rec[:, 0:3].T.copy(), "density": rec[:, 3:6].T.copy(), "quantile": rec[:, 6].copy(),
            "smooth": rec[:, 7].copy(), "noisePrev": rec[:, 8].copy(), "magn": rec[:, 9].copy(), "logLrt": rec[:, 10].copy(),
            "magnAvgPause": rec[:, 11].copy(), "featureData": st[16:23].copy(), "priorModelPars": st[8:15].copy(),
            "prior": float(st[15])}


def bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a, np.float32).view(np.uint32), np.ascontiguousarray(b, np.float32).view(np.uint32))
