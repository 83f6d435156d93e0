"""The float suppressor's DECISION STATE on the GPU is the reference's, bit for bit.

The algorithm branches on comparisons decided by the last bit of log|X[k]| (lmagn > lquantile, |lmagn -
lquantile| < WIDTH: ns_core.c:243,252 -- 774 per frame); one that goes the other way moves a quantile by a
whole tracker step and the output leaves the 1e-4 FS / 90 dB tolerance for seconds.  The kernel therefore
reproduces the reference's forward FFT rounding for rounding and takes the tracker's logarithm as the
rounded double-precision one (csrc/ns_warp.cuh ooura_fwd, nsb_log_rn).  What that buys is checked here:
after thousands of frames the three log-quantile trackers, their densities and the magnitude spectrum of
the last frame are IDENTICAL to the fields of the reference's struct, and the continuous quantities
(noise, gains, probabilities) agree to a few ulp."""
import numpy as np
import pytest

from conftest import bits_equal, gpu_nsf_state, ref_ns_trace

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fs,mode,frames,chunks", [(16000, 2, 2500, [1, 2, 7, 40, 250, 10 ** 9]), (8000, 3, 1500, [100]),
                                                  (16000, 0, 700, [10 ** 9])])
def test_tracker_state_is_the_reference_struct(nslib, reflib, fs, mode, frames, chunks):
    n, fl = 16, fs // 100
    x = nslib.synth_pcm_host(n, fs, frames * fl, base_seed=31337)
    b = nslib.NsBatch(n, fs, mode)
    xin = x.astype(np.float32).reshape(n, frames, 1, fl)
    f0, ci = 0, 0
    while f0 < frames:
        nf = min(chunks[min(ci, len(chunks) - 1)], frames - f0)
        b.process_bands_f32(np.ascontiguousarray(xin[:, f0:f0 + nf]))
        f0 += nf
        ci += 1
    lib = nslib.load_library()
    worst = {}
    for s in range(n):
        _, tr = ref_ns_trace(reflib, fs, mode, x[s])
        g = gpu_nsf_state(lib, b._handles[s], fs)
        for key in ("lquantile", "density", "magn"):
            assert bits_equal(g[key], tr[key][-1]), "stream %d: %s differs from the reference struct (max |d| %g)" % (
                s, key, np.abs(g[key] - tr[key][-1]).max())
        # continuous quantities: a few ulp (single-precision exp / tanh / tree sums on the device)
        for key in ("quantile", "smooth", "noisePrev", "logLrt", "magnAvgPause"):
            ref = tr[key][-1]
            rel = float(np.max(np.abs(g[key] - ref) / np.maximum(np.abs(ref), 1e-3)))
            worst[key] = max(worst.get(key, 0.0), rel)
        assert abs(g["prior"] - float(tr["prior"][-1])) <= 1e-5
    print("worst relative deviation of the continuous state:", {k: "%.2e" % v for k, v in worst.items()})
    assert max(worst.values()) <= 1e-4
    b.close()
