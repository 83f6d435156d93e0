"""The author's wrapper semantics (APM_NS::processCaptureStream, libapm/src/apm_ns.cpp): interleaved
multi-channel int16 / float blocks processed in place, against the reference class compiled into
oracle/_ref.  Float NS inside, hence the float parity gates."""
import ctypes as C

import numpy as np
import pytest

from conftest import _ptr, judge_float, summarize_parity

pytestmark = pytest.mark.gpu


def _ref_run(reflib, fs, mode, x_il, channels, as_float):
    lib = reflib.lib
    lib.ref_apm_ns_create.restype = C.c_void_p
    fl = fs // 100
    a = C.c_void_p(lib.ref_apm_ns_create(fs, mode, fl, channels))
    assert a.value
    out = x_il.copy()
    frames = out.shape[0] // (fl * channels)
    for f in range(frames):
        blk = np.ascontiguousarray(out[f * fl * channels:(f + 1) * fl * channels])
        if as_float:
            lib.ref_apm_ns_process_f32(a, _ptr(blk), fl, channels)
        else:
            lib.ref_apm_ns_process_i16(a, _ptr(blk), fl, channels)
        out[f * fl * channels:(f + 1) * fl * channels] = blk
    lib.ref_apm_ns_free(a)
    return out


@pytest.mark.parametrize("fs,channels,as_float", [(16000, 2, False), (48000, 2, False), (16000, 1, True), (32000, 2, True)])  # the reference AudioBuffer asserts <= 2 channels
def test_interleaved_matches_apm_ns(nslib, reflib, fs, channels, as_float):
    mode, frames, fl = 2, 400, fs // 100
    lib = nslib.load_library()
    x = nslib.synth_pcm_host(channels, fs, frames * fl, first_stream=1)      # planar [ch][n]
    il16 = np.ascontiguousarray(x.T.reshape(-1))                             # [n][ch]
    if as_float:
        il = (il16.astype(np.float32) / 32768.0).astype(np.float32)
    else:
        il = il16
    ref = _ref_run(reflib, fs, mode, il, channels, as_float)
    hs = (C.c_void_p * channels)()
    for c in range(channels):
        h = C.c_void_p()
        assert lib.WebRtcNs_Create(C.byref(h)) == 0
        hs[c] = h
    assert lib.WebRtcNs_InitBatch(hs, channels, fs, mode) == 0
    out = il.copy()
    # ragged blocks of whole frames, in place
    pos = 0
    for nf in [1, 3, 50, 10 ** 9]:
        nf = min(nf, frames - pos)
        if nf <= 0:
            break
        blk = np.ascontiguousarray(out[pos * fl * channels:(pos + nf) * fl * channels])
        fn = lib.WebRtcNs_ProcessInterleavedF32 if as_float else lib.WebRtcNs_ProcessInterleavedI16
        assert fn(hs, channels, _ptr(blk), nf * fl) == 0, lib.WebRtcNsB200_LastError()
        out[pos * fl * channels:(pos + nf) * fl * channels] = blk
        pos += nf
    for c in range(channels):
        lib.WebRtcNs_Free(hs[c])
    scale = 32768.0 if as_float else 1.0
    res = [judge_float(ref[c::channels].astype(np.float64) * scale, out[c::channels].astype(np.float64) * scale,
                       slack=1.0) for c in range(channels)]
    summarize_parity(res, "interleaved fs=%d ch=%d float=%s" % (fs, channels, as_float), 1.0)
