"""Constant tables: ours are built by closed forms (oracle/nsx_oracle.c, csrc/nsx_host_init.h,
csrc/nsf_host_init.h); when the reference tree is present they are compared entry by entry with the
literal tables in its sources.  Without the tree the oracle's and the product's tables are still
compared with each other."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, _ptr

REF = "/root/reference/WebRtc_AMP_Port/webrtc"
NAMES = ["kSinTable1024", "kBlocks160w256x", "kBlocks80w128x", "WebRtcNsx_kLogTable", "WebRtcNsx_kCounterDiv",
         "WebRtcNsx_kLogTableFrac", "kFactor1Table", "kFactor2Aggressiveness1", "kFactor2Aggressiveness2",
         "kFactor2Aggressiveness3", "kSumLogIndex", "kSumSquareLogIndex", "kLogIndex", "kDeterminantEstMatrix",
         "kIndicatorTable"]


@pytest.fixture(scope="module")
def oracle():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle_ns.so"))
    lib.nsx_oracle_table.restype = C.POINTER(C.c_int16)
    return lib


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "simt_emu")], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    return C.CDLL(os.path.join(ROOT, "tests", "simt_emu", "_build", "libns_emu.so"))


def oracle_table(lib, name):
    n = C.c_int(0)
    p = lib.nsx_oracle_table(name.encode(), C.byref(n))
    return np.array([p[i] for i in range(n.value)], np.int64)


def ref_table(name, kind=r"-?\d+"):
    srcs = ["modules/audio_processing/ns/nsx_core.c", "modules/audio_processing/ns/nsx_core_c.c",
            "common_audio/signal_processing/complex_fft_tables.h", "modules/audio_processing/ns/windows_private.h"]
    for s in srcs:
        txt = open(os.path.join(REF, s)).read()
        m = re.search(re.escape(name) + r"\[\d*\]\s*=\s*\{(.*?)\};", txt, re.S)
        if m:
            return re.findall(kind, re.sub(r"\(float\)", "", m.group(1)))
    raise KeyError(name)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
@pytest.mark.parametrize("name", NAMES)
def test_oracle_tables_equal_reference_literals(oracle, name):
    ours = oracle_table(oracle, name)
    ref = np.array([int(v) for v in ref_table(name)], np.int64)
    lo = 1 if name in ("kSumLogIndex", "kSumSquareLogIndex", "kDeterminantEstMatrix", "kLogIndex") else 0
    assert len(ours) >= len(ref)
    assert np.array_equal(ours[lo:len(ref)], ref[lo:])     # entry 0 of the log tables is "invalid" there


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_float_windows_equal_reference_literals(emu):
    w256 = np.zeros(256, np.float32)
    w128 = np.zeros(128, np.float32)
    emu.emu_nsf_tables(_ptr(w256), _ptr(w128))
    for name, ours in (("kBlocks160w256", w256), ("kBlocks80w128", w128)):
        ref = np.array([np.float32(float(v)) for v in ref_table(name, r"\d+\.\d+")], np.float32)
        assert len(ref) == len(ours)
        assert np.array_equal(ref.view(np.uint32), ours.view(np.uint32)), name   # bit for bit


def test_product_tables_equal_oracle_tables(oracle, emu):
    a = {k: np.zeros(n, np.int16) for k, n in (("w256", 256), ("w128", 128), ("lf", 256), ("cd", 201), ("lt", 9),
                                               ("li", 129), ("f1", 257), ("f2", 3 * 257), ("ind", 17), ("m", 5))}
    tw = np.zeros(128, np.uint32)
    emu.emu_nsx_tables(_ptr(a["w256"]), _ptr(a["w128"]), _ptr(tw), _ptr(a["lf"]), _ptr(a["cd"]), _ptr(a["lt"]),
                       _ptr(a["li"]), _ptr(a["f1"]), _ptr(a["f2"]), _ptr(a["ind"]), _ptr(a["m"]))
    t = lambda n: oracle_table(oracle, n)
    assert np.array_equal(a["w256"], t("kBlocks160w256x")) and np.array_equal(a["w128"], t("kBlocks80w128x"))
    assert np.array_equal(a["lf"], t("WebRtcNsx_kLogTableFrac")) and np.array_equal(a["cd"], t("WebRtcNsx_kCounterDiv"))
    assert np.array_equal(a["lt"], t("WebRtcNsx_kLogTable")) and np.array_equal(a["li"][1:], t("kLogIndex")[1:])
    assert np.array_equal(a["f1"], t("kFactor1Table")) and np.array_equal(a["ind"], t("kIndicatorTable"))
    for k in range(3):
        assert np.array_equal(a["f2"][257 * k:257 * (k + 1)], t("kFactor2Aggressiveness%d" % (k + 1)))
    sin = t("kSinTable1024")
    assert np.array_equal((tw & 0xffff).astype(np.uint16).view(np.int16), sin[4 * np.arange(128) + 256])
    assert np.array_equal((tw >> 16).astype(np.uint16).view(np.int16), sin[4 * np.arange(128)])
    assert a["m"][0] == t("kSumLogIndex")[5] and a["m"][1] == t("kSumSquareLogIndex")[5]
    assert a["m"][2] == t("kDeterminantEstMatrix")[5]
    assert a["m"][3] == t("kSumLogIndex")[65] and a["m"][4] == t("kSumSquareLogIndex")[65]
