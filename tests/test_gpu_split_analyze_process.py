"""Analyze and Process fed different signals on the GPU against the compiled
reference driven through its own API: the band-frame float entry, the int16 PCM entries (host and
device pointers), the single-stream WebRtcNs_Analyze / WebRtcNs_Process pair, and the hand-over of
a running stream from the fused kernel."""
import ctypes as C

import numpy as np
import pytest

from conftest import FLOAT_MAX_ABS, FLOAT_MIN_SNR, judge_float, summarize_parity
from test_emulated_kernels import split_signals

pytestmark = pytest.mark.gpu


def _handles(lib, n, fs, mode):
    hs = (C.c_void_p * n)()
    for i in range(n):
        h = C.c_void_p()
        assert lib.WebRtcNs_Create(C.byref(h)) == 0, lib.WebRtcNsB200_LastError()
        hs[i] = h
    assert lib.WebRtcNs_InitBatch(hs, n, fs, mode) == 0, lib.WebRtcNsB200_LastError()
    return hs


def _free(lib, hs):
    for h in hs:
        lib.WebRtcNs_Free(h)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("fs,mode,nb,fused", [(16000, 2, 1, 0), (16000, 1, 3, 0), (32000, 2, 2, 25), (8000, 3, 1, 10)])
def test_band_frame_entry(nslib, reflib, fs, mode, nb, fused):
    lib = nslib.load_library()
    n, frames = 8, 300
    band_fs = 8000 if fs == 8000 else 16000
    fl = band_fs // 100
    ana, x = split_signals(nslib.synth_pcm_host, n, band_fs, frames, nb)
    hs = _handles(lib, n, fs, mode)
    out = np.zeros_like(x)
    per = frames * nb * fl
    if fused:
        # a stream that ran fused first: the split kernel picks its state up
        xa = np.ascontiguousarray(x[:, :fused])
        oa = np.zeros_like(xa)
        assert lib.WebRtcNs_ProcessBatchBandsF32(hs, n, nb, _p(xa), fused * nb * fl, _p(oa), fused * nb * fl, fused) == 0
        out[:, :fused] = oa
    xb = np.ascontiguousarray(x[:, fused:])
    ab = np.ascontiguousarray(ana.reshape(n, frames, fl)[:, fused:])
    ob = np.zeros_like(xb)
    rest = frames - fused
    assert lib.WebRtcNs_AnalyzeProcessBatchBandsF32(hs, n, nb, _p(ab), rest * fl, _p(xb), rest * nb * fl, _p(ob),
                                                    rest * nb * fl, rest) == 0, lib.WebRtcNsB200_LastError()
    out[:, fused:] = ob
    res = []
    for s in range(n):
        ref = reflib.ns_split(fs, mode, ana[s], x[s], nb, fused)
        res.append(judge_float(ref.ravel(), out[s].ravel()))
    summarize_parity(res, "split bands fs=%d nb=%d" % (fs, nb), 1.0, max_abs=FLOAT_MAX_ABS, min_snr=FLOAT_MIN_SNR)
    _free(lib, hs)
    assert per > 0


@pytest.mark.parametrize("device_ptrs", [False, True])
def test_int16_pcm_entries(nslib, reflib, device_ptrs):
    import torch
    lib = nslib.load_library()
    fs, mode, n, frames, fl = 16000, 2, 8, 250, 160
    anaf, xf = split_signals(nslib.synth_pcm_host, n, fs, frames, 1)
    ana = np.ascontiguousarray(np.clip(np.rint(anaf), -32768, 32767).astype(np.int16))
    x = np.ascontiguousarray(xf.reshape(n, frames * fl).astype(np.int16))
    hs = _handles(lib, n, fs, mode)
    out = np.zeros_like(x)
    # ragged launches: state (both histories, the four split arrays) round-trips HBM
    f0 = 0
    for nf in [1, 7, 100, 10 ** 9]:
        nf = min(nf, frames - f0)
        if nf <= 0:
            break
        a = np.ascontiguousarray(ana[:, f0 * fl:(f0 + nf) * fl])
        b = np.ascontiguousarray(x[:, f0 * fl:(f0 + nf) * fl])
        o = np.zeros_like(b)
        if device_ptrs:
            ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
            to = torch.empty_like(tb)
            assert lib.WebRtcNs_AnalyzeProcessBatchDevice(hs, n, C.c_void_p(ta.data_ptr()), nf * fl, C.c_void_p(tb.data_ptr()),
                                                          nf * fl, C.c_void_p(to.data_ptr()), nf * fl, nf, None) == 0
            assert lib.WebRtcNsB200_Synchronize() == 0
            o = to.cpu().numpy()
        else:
            assert lib.WebRtcNs_AnalyzeProcessBatch(hs, n, _p(a), nf * fl, _p(b), nf * fl, _p(o), nf * fl, nf) == 0, (
                lib.WebRtcNsB200_LastError())
        out[:, f0 * fl:(f0 + nf) * fl] = o
        f0 += nf
    res = []
    for s in range(n):
        ref = reflib.ns_split(fs, mode, ana[s].astype(np.float32), x[s].astype(np.float32).reshape(frames, 1, fl), 1, 0)
        refi = np.clip(np.where(ref > 0, np.floor(ref + 0.5), np.ceil(ref - 0.5)), -32768, 32767).ravel()
        res.append(judge_float(refi, out[s].astype(np.float32), slack=1.0))
    summarize_parity(res, "split int16 pcm", 1.0, max_abs=1.0)
    _free(lib, hs)


def test_split_kernel_on_one_signal_equals_fused(nslib):
    """Same signal to both sides: the split kernel (two FFTs, separate memories) must reproduce
    the fused kernel bit for bit -- and a stream in split mode keeps doing so on fused calls."""
    lib = nslib.load_library()
    fs, mode, n, frames, fl = 16000, 2, 8, 300, 160
    x = nslib.synth_pcm_host(n, fs, frames * fl)
    a = _handles(lib, n, fs, mode)
    b = _handles(lib, n, fs, mode)
    oa, ob = np.zeros_like(x), np.zeros_like(x)
    assert lib.WebRtcNs_ProcessBatch(a, n, _p(x), frames * fl, _p(oa), frames * fl, frames) == 0
    half = frames // 2
    x1, x2 = np.ascontiguousarray(x[:, :half * fl]), np.ascontiguousarray(x[:, half * fl:])
    o1, o2 = np.zeros_like(x1), np.zeros_like(x2)
    assert lib.WebRtcNs_AnalyzeProcessBatch(b, n, _p(x1), half * fl, _p(x1), half * fl, _p(o1), half * fl, half) == 0
    assert lib.WebRtcNs_ProcessBatch(b, n, _p(x2), (frames - half) * fl, _p(o2), (frames - half) * fl, frames - half) == 0
    ob = np.concatenate([o1, o2], axis=1)
    assert np.array_equal(oa, ob)
    _free(lib, a)
    _free(lib, b)


def test_single_stream_api_with_distinct_frames(nslib, reflib):
    lib = nslib.load_library()
    fs, mode, frames, fl = 16000, 2, 60, 160
    ana, x = split_signals(nslib.synth_pcm_host, 1, fs, frames, 1)
    h = C.c_void_p()
    assert lib.WebRtcNs_Create(C.byref(h)) == 0
    assert lib.WebRtcNs_Init(h, fs) == 0 and lib.WebRtcNs_set_policy(h, mode) == 0
    out = np.zeros(frames * fl, np.float32)
    a = ana[0].reshape(frames, fl)
    for f in range(frames):
        fin = np.ascontiguousarray(x[0, f, 0])
        fo = np.zeros(fl, np.float32)
        pin = (C.c_void_p * 1)(fin.ctypes.data)
        pout = (C.c_void_p * 1)(fo.ctypes.data)
        lib.WebRtcNs_Analyze(h, _p(np.ascontiguousarray(a[f])))
        lib.WebRtcNs_Process(h, pin, 1, pout)
        out[f * fl:(f + 1) * fl] = fo
    ref = reflib.ns_split(fs, mode, ana[0], x[0], 1, 0).ravel()
    r = judge_float(ref, out)
    assert r[1], "outside envelope: %.3f LSB %.1f dB" % (r[2], r[3])
    lib.WebRtcNs_Free(h)


def test_unsupported_combinations_fail_loudly(nslib):
    lib = nslib.load_library()
    hs = _handles(lib, 1, 48000, 2)
    x = np.zeros(480, np.int16)
    o = np.zeros(480, np.int16)
    assert lib.WebRtcNs_AnalyzeProcessBatch(hs, 1, _p(x), 480, _p(x), 480, _p(o), 480, 1) == -1
    assert b"8/16 kHz" in lib.WebRtcNsB200_LastError()
    assert lib.WebRtcNs_AnalyzeProcessBatch(hs, 1, None, 480, _p(x), 480, _p(o), 480, 1) == -1
    _free(lib, hs)
