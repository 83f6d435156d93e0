"""Several GPUs inside ONE process (the batch call buckets handles by device and runs the
devices side by side; 8b "Threading": independent handles on independent threads).  Needs >= 2 visible
GPUs -- run with `gpurun --gpus 2` (the log of such a run is committed under profiles/); skipped on one."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _need_two(nslib):
    lib = nslib.load_library()
    if lib.WebRtcNsB200_DeviceCount() < 2:
        pytest.skip("one GPU")
    return lib


def _batch_on(nslib, placement, fs, mode, fixed):
    """One NsBatch whose stream s lives on GPU placement[s] (any interleaving)."""
    import ctypes as C
    lib = nslib.load_library()
    b = nslib.NsBatch.__new__(nslib.NsBatch)
    b._lib, b.n, b.fs, b.fixed = lib, len(placement), fs, fixed
    b._p = "WebRtcNsx" if fixed else "WebRtcNs"
    b._handles = (C.c_void_p * b.n)()
    for s, dev in enumerate(placement):
        assert lib.WebRtcNsB200_SetCreateDevice(dev) == 0
        h = C.c_void_p()
        assert getattr(lib, b._p + "_Create")(C.byref(h)) == 0
        b._handles[s] = h
    lib.WebRtcNsB200_SetCreateDevice(-1)
    assert getattr(lib, b._p + "_InitBatch")(b._handles, b.n, fs, mode) == 0, lib.WebRtcNsB200_LastError()
    return b


@pytest.mark.parametrize("fs", [16000, 48000])
def test_one_call_spanning_devices_in_any_order(nslib, reflib, fs):
    """Handles of two GPUs interleaved unevenly in one list ([0,0,0,0,0,1,0,0,0,1,1,0,...]): every device gets
    ONE launch over all of its streams, whatever their positions in the list; bit-exact per stream (NSx)."""
    _need_two(nslib)
    mode, frames = 2, 130
    fl = fs // 100
    placement = [0, 0, 0, 0, 0, 1, 0, 0, 0, 1, 1, 0, 1, 1, 1, 1, 0, 1, 0]
    n = len(placement)
    x = nslib.synth_pcm_host(n, fs, frames * fl)
    b = _batch_on(nslib, placement, fs, mode, True)
    out = np.zeros_like(x)
    for f0, nf in ((0, 7), (7, 100), (107, 23)):     # several calls: the staging pipeline is reused and re-cut
        out[:, f0 * fl:(f0 + nf) * fl] = b.process(np.ascontiguousarray(x[:, f0 * fl:(f0 + nf) * fl]))
    for s in range(n):
        assert np.array_equal(out[s], reflib.nsx(fs, mode, x[s])), "stream %d on GPU %d" % (s, placement[s])
    b.close()


def test_full_size_batch_over_two_devices_matches_per_device_runs(nslib, reflib):
    """4096 NSx streams, half per GPU, 100 frames in one WebRtcNsx_ProcessBatch call == two one-GPU batches."""
    _need_two(nslib)
    fs, mode, frames, n = 16000, 2, 100, 4096
    x = nslib.synth_pcm_host(n, fs, frames * 160)
    both = nslib.NsBatch(n, fs, mode, fixed=True, devices=[0, 1])
    out = both.process(x)
    both.close()
    for dev, sl in ((0, slice(0, n // 2)), (1, slice(n // 2, n))):
        one = nslib.NsBatch(n // 2, fs, mode, fixed=True, devices=[dev])
        assert np.array_equal(out[sl], one.process(np.ascontiguousarray(x[sl])))
        one.close()
    for s in (0, n // 2 - 1, n // 2, n - 1):
        assert np.array_equal(out[s], reflib.nsx(fs, mode, x[s]))


def test_two_threads_two_devices_run_side_by_side(nslib, reflib):
    """One host thread per GPU, each ticking its own batch: correct, and concurrent -- the two loops together
    take clearly less than the sum of each alone (no process-wide lock around the GPU waits)."""
    import time
    _need_two(nslib)
    fs, mode, n, frames, reps = 16000, 2, 2048, 100, 12
    xs = [nslib.synth_pcm_host(n, fs, frames * 160, base_seed=7 + d) for d in range(2)]
    bs = [nslib.NsBatch(n, fs, mode, fixed=True, devices=[d]) for d in range(2)]
    outs = [None, None]

    def loop(d):
        for _ in range(reps):
            outs[d] = bs[d].process(xs[d])

    for d in range(2):
        loop(d)                     # warm-up (allocations, first-touch)
    t0 = time.perf_counter()
    loop(0)
    t_a = time.perf_counter() - t0
    t0 = time.perf_counter()
    loop(1)
    t_b = time.perf_counter() - t0
    th = [threading.Thread(target=loop, args=(d,)) for d in range(2)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    t_both = time.perf_counter() - t0
    print("alone %.3f + %.3f s, together %.3f s" % (t_a, t_b, t_both))
    assert t_both < 0.8 * (t_a + t_b)
    # state carried across all the calls above: compare with the reference on the concatenated input
    for d in range(2):
        whole = np.concatenate([xs[d][0]] * (3 * reps))
        ref = reflib.nsx(fs, mode, whole)
        assert np.array_equal(outs[d][0], ref[-frames * 160:])
    for b in bs:
        b.close()
