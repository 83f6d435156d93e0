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
