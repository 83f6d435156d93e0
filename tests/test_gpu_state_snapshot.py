"""Stream state snapshot / restore / migration: a stream exported after some
frames and imported into a fresh handle continues bit-identically to the uninterrupted stream --
for the fixed-point suppressor that also means bit-identical to the compiled reference."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make(lib, fixed, fs, mode, n):
    hs = (C.c_void_p * n)()
    create = lib.WebRtcNsx_Create if fixed else lib.WebRtcNs_Create
    for i in range(n):
        h = C.c_void_p()
        assert create(C.byref(h)) == 0, lib.WebRtcNsB200_LastError()
        hs[i] = h
    init = lib.WebRtcNsx_InitBatch if fixed else lib.WebRtcNs_InitBatch
    assert init(hs, n, fs, mode) == 0, lib.WebRtcNsB200_LastError()
    return hs


def _run(lib, fixed, hs, n, x, frames):
    x = np.ascontiguousarray(x)
    out = np.zeros_like(x)
    fn = lib.WebRtcNsx_ProcessBatch if fixed else lib.WebRtcNs_ProcessBatch
    assert fn(hs, n, x.ctypes.data_as(C.c_void_p), x.shape[1], out.ctypes.data_as(C.c_void_p), out.shape[1],
              frames) == 0, lib.WebRtcNsB200_LastError()
    return out


@pytest.mark.parametrize("fixed,fs", [(True, 16000), (True, 48000), (False, 16000), (False, 32000)])
def test_export_import_continues_bit_identically(nslib, reflib, fixed, fs):
    lib = nslib.load_library()
    n, mode, fl, head, total = 4, 2, fs // 100, 130, 260
    x = nslib.synth_pcm_host(n, fs, total * fl)
    free = lib.WebRtcNsx_Free if fixed else lib.WebRtcNs_Free
    # uninterrupted
    a = _make(lib, fixed, fs, mode, n)
    whole = np.concatenate([_run(lib, fixed, a, n, x[:, :head * fl], head),
                            _run(lib, fixed, a, n, x[:, head * fl:], total - head)], axis=1)
    # interrupted: export after `head` frames, import into brand-new handles (never initialised)
    b = _make(lib, fixed, fs, mode, n)
    first = _run(lib, fixed, b, n, x[:, :head * fl], head)
    blobs = []
    for i in range(n):
        size = lib.WebRtcNsB200_StateSize(b[i])
        assert size > 0
        buf = (C.c_ubyte * size)()
        assert lib.WebRtcNsB200_ExportState(b[i], buf, size) == 0, lib.WebRtcNsB200_LastError()
        blobs.append(buf)
        free(b[i])
    c = (C.c_void_p * n)()
    create = lib.WebRtcNsx_Create if fixed else lib.WebRtcNs_Create
    for i in range(n):
        h = C.c_void_p()
        assert create(C.byref(h)) == 0
        c[i] = h
        assert lib.WebRtcNsB200_ImportState(h, blobs[i], len(blobs[i])) == 0, lib.WebRtcNsB200_LastError()
    second = _run(lib, fixed, c, n, x[:, head * fl:], total - head)
    resumed = np.concatenate([first, second], axis=1)
    assert np.array_equal(whole, resumed)
    if fixed:
        for s in range(n):
            assert np.array_equal(resumed[s], reflib.nsx(fs, mode, x[s]))
    for i in range(n):
        free(a[i])
        free(c[i])


def test_import_rejects_wrong_kind_and_garbage(nslib):
    lib = nslib.load_library()
    f = _make(lib, False, 16000, 1, 1)
    xh = _make(lib, True, 16000, 1, 1)
    size = lib.WebRtcNsB200_StateSize(f[0])
    buf = (C.c_ubyte * size)()
    assert lib.WebRtcNsB200_ExportState(f[0], buf, size) == 0
    assert lib.WebRtcNsB200_ImportState(xh[0], buf, size) == -1          # float blob into a fixed handle
    assert lib.WebRtcNsB200_ImportState(f[0], buf, size - 1) == -1        # truncated
    junk = (C.c_ubyte * size)()
    assert lib.WebRtcNsB200_ImportState(f[0], junk, size) == -1           # no tag
    assert lib.WebRtcNsB200_ExportState(f[0], buf, 16) == -1              # buffer too small
    # control words out of range: fs, a tracker counter (header 696 bytes, then the slab: counters at words 2-4)
    bad = (C.c_ubyte * size).from_buffer_copy(bytes(buf))
    bad[8:12] = (44100).to_bytes(4, "little")
    assert lib.WebRtcNsB200_ImportState(f[0], bad, size) == -1
    bad = (C.c_ubyte * size).from_buffer_copy(bytes(buf))
    bad[696 + 8:696 + 12] = (100000).to_bytes(4, "little")
    assert lib.WebRtcNsB200_ImportState(f[0], bad, size) == -1
    assert b"out of range" in lib.WebRtcNsB200_LastError()
    assert lib.WebRtcNsB200_ImportState(f[0], buf, size) == 0             # the untouched blob still goes in
    assert lib.WebRtcNsB200_StateSize(None) == 0
    lib.WebRtcNs_Free(f[0])
    lib.WebRtcNsx_Free(xh[0])


def test_migrate_between_gpus(nslib, reflib):
    """Needs two GPUs: a fixed-point stream hops to GPU 1 mid-stream and stays bit-exact."""
    lib = nslib.load_library()
    if lib.WebRtcNsB200_DeviceCount() < 2:
        pytest.skip("one GPU")
    fs, mode, fl, head, total = 16000, 2, 160, 77, 200
    x = nslib.synth_pcm_host(2, fs, total * fl)
    hs = _make(lib, True, fs, mode, 2)
    first = _run(lib, True, hs, 2, x[:, :head * fl], head)
    assert lib.WebRtcNsB200_MigrateHandle(hs[1], 1) == 0, lib.WebRtcNsB200_LastError()
    assert lib.WebRtcNsB200_HandleDevice(hs[1]) == 1 and lib.WebRtcNsB200_HandleDevice(hs[0]) == 0
    second = _run(lib, True, hs, 2, x[:, head * fl:], total - head)   # one batch spanning both GPUs
    out = np.concatenate([first, second], axis=1)
    for s in range(2):
        assert np.array_equal(out[s], reflib.nsx(fs, mode, x[s]))
    for i in range(2):
        lib.WebRtcNsx_Free(hs[i])
