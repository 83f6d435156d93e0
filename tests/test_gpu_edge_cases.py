"""Edge cases of the batched boundary, checked bit for bit through the fixed-point suppressor (and
through the float one where the contract differs): empty calls, the smallest and odd batch sizes,
ragged and padded strides, in-place buffers (the reference's callers pass the same buffers in and
out: libapm/src/apm_ns.cpp:71-73, ns_core.c:1225-1235), all-zero and full-scale input, a very long
single launch, and handle lists that change between calls (the library remembers the last
validated list)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_zero_frames_is_a_no_op(nslib):
    n, fs = 4, 16000
    b = nslib.NsBatch(n, fs, 2, fixed=True)
    x = np.zeros((n, 160), np.int16)
    out = np.full((n, 160), 77, np.int16)
    b.process_ptr(x.ctypes.data, 160, out.ctypes.data, 160, 0)
    assert (out == 77).all()
    b.close()


@pytest.mark.parametrize("n", [1, 3, 33, 65, 149])
@pytest.mark.parametrize("fixed", [True, False])
def test_odd_batch_sizes(nslib, reflib, n, fixed):
    """Batches that leave CTAs / warps partly empty."""
    fs, mode, frames = 16000, 2, 60
    x = nslib.synth_pcm_host(n, fs, frames * 160)
    b = nslib.NsBatch(n, fs, mode, fixed=fixed)
    out = b.process(x)
    for s in sorted({0, n // 2, n - 1}):
        if fixed:
            assert np.array_equal(out[s], reflib.nsx(fs, mode, x[s]))
        else:
            _, refi, _ = reflib.ns(fs, mode, x[s])
            assert np.abs(out[s].astype(np.int32) - refi.astype(np.int32)).max() <= 1
    b.close()


def test_padded_and_unequal_strides_leave_padding_alone(nslib, reflib):
    fs, mode, frames, n = 8000, 1, 50, 5
    fl = fs // 100
    x = nslib.synth_pcm_host(n, fs, frames * fl)
    xin = np.full((n, frames * fl + 14), 1234, np.int16)
    xin[:, :frames * fl] = x
    out = np.full((n, frames * fl + 102), -4321, np.int16)
    b = nslib.NsBatch(n, fs, mode, fixed=True)
    b.process_ptr(xin.ctypes.data, xin.shape[1], out.ctypes.data, out.shape[1], frames)
    for s in range(n):
        assert np.array_equal(out[s, :frames * fl], reflib.nsx(fs, mode, x[s]))
    assert (out[:, frames * fl:] == -4321).all()
    assert (xin[:, frames * fl:] == 1234).all() and np.array_equal(xin[:, :frames * fl], x)
    b.close()


@pytest.mark.parametrize("fs", [16000, 48000])
def test_in_place_buffers(nslib, reflib, fs):
    """pcm_out == pcm_in, host pointers and device pointers."""
    import torch
    mode, frames, n = 2, 40, 6
    fl = fs // 100
    x = nslib.synth_pcm_host(n, fs, frames * fl)
    want = np.stack([reflib.nsx(fs, mode, x[s]) for s in range(n)])
    b = nslib.NsBatch(n, fs, mode, fixed=True)
    buf = x.copy()
    b.process_ptr(buf.ctypes.data, buf.shape[1], buf.ctypes.data, buf.shape[1], frames)
    assert np.array_equal(buf, want)
    b.reset(mode)
    d = torch.from_numpy(x.copy()).cuda()
    b.process_device(d.data_ptr(), d.shape[1], d.data_ptr(), d.shape[1], frames)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), want)
    b.close()


def test_float_single_stream_in_place(nslib, reflib):
    """WebRtcNs_Process with outframe == spframe, as both reference wrappers call it."""
    fs, mode, frames = 16000, 2, 30
    x = nslib.synth_pcm_host(3, fs, frames * 160)[2]
    ref, _, _ = reflib.ns(fs, mode, x)
    lib = nslib.load_library()
    h = C.c_void_p()
    assert lib.WebRtcNs_Create(C.byref(h)) == 0 and lib.WebRtcNs_Init(h, fs) == 0 and lib.WebRtcNs_set_policy(h, mode) == 0
    got = np.zeros(frames * 160, np.float32)
    for f in range(frames):
        fr = x[f * 160:(f + 1) * 160].astype(np.float32)
        p = (C.c_void_p * 1)(fr.ctypes.data)
        lib.WebRtcNs_Analyze(h, C.c_void_p(fr.ctypes.data))
        lib.WebRtcNs_Process(h, p, 1, p)
        got[f * 160:(f + 1) * 160] = fr
    lib.WebRtcNs_Free(h)
    assert np.abs(got - ref).max() <= 3.2768      # 1e-4 full scale


@pytest.mark.parametrize("fixed", [True, False])
def test_all_zero_then_full_scale(nslib, reflib, fixed):
    """Digital silence from t0 (zero-energy path, start-up counted in non-silent frames), then a
    full-scale +-32768/32767 square wave (saturating synthesis)."""
    fs, mode = 16000, 3
    x = np.zeros(160 * 140, np.int16)
    t = np.arange(160 * 100)
    x[160 * 40:] = np.where((t // 20) % 2 == 0, 32767, -32768).astype(np.int16)
    b = nslib.NsBatch(1, fs, mode, fixed=fixed)
    out = b.process(x[None, :])[0]
    assert (out[:160 * 39] == 0).all()
    if fixed:
        assert np.array_equal(out, reflib.nsx(fs, mode, x))
    else:
        _, refi, _ = reflib.ns(fs, mode, x)
        assert np.abs(out.astype(np.int32) - refi.astype(np.int32)).max() <= 3
    b.close()


def test_one_long_launch_equals_the_reference(nslib, reflib):
    """A single stream, 12 000 frames (2 minutes) in ONE launch: the frame loop, the prefetch and the
    threshold re-estimation (every 512 frames) with no launch boundary to reset anything."""
    fs, mode, frames = 16000, 2, 12000
    x = nslib.synth_pcm_host(4, fs, frames * 160)[3]
    b = nslib.NsBatch(1, fs, mode, fixed=True)
    assert np.array_equal(b.process(x[None, :])[0], reflib.nsx(fs, mode, x))
    b.close()


def test_changing_handle_lists_between_calls(nslib, reflib):
    """The validated handle list is remembered between calls: a freed and re-created batch (the
    allocator may hand out the same addresses), a sub-list, a permuted list and a re-Init at another
    rate must all be seen."""
    mode, frames = 2, 30
    x16 = nslib.synth_pcm_host(8, 16000, frames * 160)
    x8 = nslib.synth_pcm_host(8, 8000, frames * 80)
    for _ in range(3):
        a = nslib.NsBatch(8, 16000, mode, fixed=True)
        o = a.process(x16)
        assert np.array_equal(o[5], reflib.nsx(16000, mode, x16[5]))
        a.close()
        c = nslib.NsBatch(8, 8000, mode, fixed=True)
        o = c.process(x8)
        assert np.array_equal(o[5], reflib.nsx(8000, mode, x8[5]))
        c.close()
    b = nslib.NsBatch(8, 16000, mode, fixed=True)
    lib = nslib.load_library()
    half = frames // 2 * 160
    out = np.zeros_like(x16)
    out[:, :half] = b.process(np.ascontiguousarray(x16[:, :half]))
    # second half through a permuted handle list (and permuted rows)
    perm = [3, 0, 7, 1, 6, 2, 5, 4]
    hv = (C.c_void_p * 8)(*[b._handles[p] for p in perm])
    xin = np.ascontiguousarray(x16[perm][:, half:])
    o2 = np.zeros_like(xin)
    assert lib.WebRtcNsx_ProcessBatch(hv, 8, xin.ctypes.data_as(C.c_void_p), xin.shape[1],
                                      o2.ctypes.data_as(C.c_void_p), o2.shape[1], frames - frames // 2) == 0
    for i, p in enumerate(perm):
        out[p, half:] = o2[i]
    for s in range(8):
        assert np.array_equal(out[s], reflib.nsx(16000, mode, x16[s]))
    # a sub-list after re-Init of those handles only
    sub = (C.c_void_p * 3)(b._handles[1], b._handles[4], b._handles[6])
    assert lib.WebRtcNsx_InitBatch(sub, 3, 8000, mode) == 0
    xs = np.ascontiguousarray(x8[:3])
    os_ = np.zeros_like(xs)
    assert lib.WebRtcNsx_ProcessBatch(sub, 3, xs.ctypes.data_as(C.c_void_p), xs.shape[1],
                                      os_.ctypes.data_as(C.c_void_p), os_.shape[1], frames) == 0
    for i in range(3):
        assert np.array_equal(os_[i], reflib.nsx(8000, mode, xs[i]))
    # the full list now mixes 8 and 16 kHz handles: refused, as one batch has one frame length
    full = (C.c_void_p * 8)(*[b._handles[i] for i in range(8)])
    assert lib.WebRtcNsx_ProcessBatch(full, 8, x16.ctypes.data_as(C.c_void_p), x16.shape[1],
                                      out.ctypes.data_as(C.c_void_p), out.shape[1], 1) == -1
    b.close()


@pytest.mark.parametrize("fixed", [True, False])
def test_computed_and_listed_slots_agree(nslib, fixed):
    """Handles created one after the other sit in consecutive state slots, and the kernels then compute
    a stream's slot instead of loading it from the slot list; a permuted handle list takes the list
    path.  Both must produce the same bits, tick by tick (state round-trips HBM through the TMA
    bulk copies every call)."""
    n, fs, mode, frames = 37, 16000, 2, 12
    lib = nslib.load_library()
    x = nslib.synth_pcm_host(n, fs, frames * 160)
    a = nslib.NsBatch(n, fs, mode, fixed=fixed)
    ref = np.zeros_like(x)
    for f in range(frames):      # one frame per call
        ref[:, f * 160:(f + 1) * 160] = a.process(np.ascontiguousarray(x[:, f * 160:(f + 1) * 160]))
    a.close()
    b = nslib.NsBatch(n, fs, mode, fixed=fixed)
    perm = list(reversed(range(n)))
    hv = (C.c_void_p * n)(*[b._handles[p] for p in perm])
    fn = lib.WebRtcNsx_ProcessBatch if fixed else lib.WebRtcNs_ProcessBatch
    out = np.zeros_like(x)
    for f in range(frames):
        xin = np.ascontiguousarray(x[perm][:, f * 160:(f + 1) * 160])
        o = np.zeros_like(xin)
        assert fn(hv, n, xin.ctypes.data_as(C.c_void_p), 160, o.ctypes.data_as(C.c_void_p), 160, 1) == 0
        for i, p in enumerate(perm):
            out[p, f * 160:(f + 1) * 160] = o[i]
    b.close()
    assert ref.any()
    assert np.array_equal(ref, out)


def test_batch_validation_rejects_what_the_kernels_cannot_take(nslib):
    """Misaligned PCM pointers (the kernels move 32-bit words), a handle listed twice (two warps on one
    slab) and -- accepted -- a single stream whose stride is shorter than its frames (never used to step)."""
    lib = nslib.load_library()
    fs, frames = 16000, 10
    b = nslib.NsBatch(2, fs, 2, fixed=True)
    buf = np.zeros(2 * frames * 160 + 8, np.int16)
    out = np.zeros_like(buf)
    base, obase = buf.ctypes.data, out.ctypes.data
    assert lib.WebRtcNsx_ProcessBatch(b._handles, 2, C.c_void_p(base + 2), frames * 160, C.c_void_p(obase), frames * 160, frames) == -1
    assert b"aligned" in lib.WebRtcNsB200_LastError()
    assert lib.WebRtcNsx_ProcessBatch(b._handles, 2, C.c_void_p(base), frames * 160, C.c_void_p(obase + 2), frames * 160, frames) == -1
    twice = (C.c_void_p * 2)(b._handles[0], b._handles[0])
    assert lib.WebRtcNsx_ProcessBatch(twice, 2, C.c_void_p(base), frames * 160, C.c_void_p(obase), frames * 160, frames) == -1
    assert b"twice" in lib.WebRtcNsB200_LastError()
    one = (C.c_void_p * 1)(b._handles[1])
    x = nslib.synth_pcm_host(1, fs, frames * 160)
    o1 = np.zeros_like(x)
    assert lib.WebRtcNsx_ProcessBatch(one, 1, C.c_void_p(x.ctypes.data), 0, C.c_void_p(o1.ctypes.data), 2, frames) == 0, \
        lib.WebRtcNsB200_LastError()
    b2 = nslib.NsBatch(1, fs, 2, fixed=True)
    assert np.array_equal(o1, b2.process(x))
    b.close()
    b2.close()


def test_calls_from_several_threads(nslib, reflib):
    """Independent handles driven from independent host threads (the reference keeps no shared state,
    ns/noise_suppression.c:20-66): four threads, each with its own batch and its own create/process/free
    cycle, all at once on one device; and the last error is the calling thread's own."""
    import threading
    lib = nslib.load_library()
    fs, mode, frames, n = 16000, 2, 120, 6
    xs = [nslib.synth_pcm_host(n, fs, frames * 160, base_seed=100 + t) for t in range(4)]
    outs, errs = [None] * 4, [None] * 4

    def work(t):
        try:
            b = nslib.NsBatch(n, fs, mode, fixed=True)
            o = np.zeros_like(xs[t])
            for f0 in range(0, frames, 15):
                o[:, f0 * 160:(f0 + 15) * 160] = b.process(np.ascontiguousarray(xs[t][:, f0 * 160:(f0 + 15) * 160]))
            outs[t] = o
            if t == 1:   # provoke an error in this thread only
                assert lib.WebRtcNsx_ProcessBatch(b._handles, n, None, 1, None, 1, 1) == -1
                errs[t] = lib.WebRtcNsB200_LastError()
            else:
                errs[t] = b""
            b.close()
        except Exception as e:   # noqa: BLE001
            errs[t] = repr(e).encode()

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert errs[1] == b"strides must be even", errs
    assert errs[0] == b"" and errs[2] == b"" and errs[3] == b"", errs
    for t in range(4):
        for s in (0, n - 1):
            assert np.array_equal(outs[t][s], reflib.nsx(fs, mode, xs[t][s]))


def test_device_call_on_a_user_stream_is_ordered_before_later_library_calls(nslib, reflib):
    """A launch enqueued on the caller's stream by a *Device entry point, immediately followed by calls that
    use the library's own stream (the getter, a host-pointer batch, Free): no synchronisation by the caller."""
    import torch
    lib = nslib.load_library()
    fs, mode, n, frames = 16000, 2, 64, 300
    x = nslib.synth_pcm_host(n, fs, 2 * frames * 160)
    st = torch.cuda.Stream()
    d_in = torch.from_numpy(x[:, :frames * 160].copy()).cuda()
    d_out = torch.empty_like(d_in)
    torch.cuda.synchronize()
    b = nslib.NsBatch(n, fs, mode, fixed=True)
    b.process_device(d_in.data_ptr(), frames * 160, d_out.data_ptr(), frames * 160, frames, st.cuda_stream)
    second = b.process(np.ascontiguousarray(x[:, frames * 160:]))     # library stream, right behind it
    b.close()                                                          # Free right behind that
    torch.cuda.synchronize()
    first = d_out.cpu().numpy()
    for s in (0, 31, n - 1):
        assert np.array_equal(np.concatenate([first[s], second[s]]), reflib.nsx(fs, mode, x[s]))
