"""The float suppressor's DECISION STATE on the GPU is the reference's, bit for bit.

The algorithm branches on comparisons decided by the last bit of log|X[k]| (lmagn > lquantile, |lmagn -
lquantile| < WIDTH: ns_core.c:243,252 -- 774 per frame); one that goes the other way moves a quantile by a
whole tracker step and the output leaves the 1e-4 FS / 90 dB tolerance for seconds.  The kernel therefore
reproduces the reference's forward FFT rounding for rounding and takes the tracker's logarithm as the
rounded double-precision one (csrc/ns_warp.cuh ooura_fwd, nsb_log_rn).  What that buys is checked here:
after thousands of frames every field of the recursion -- the three log-quantile trackers and their densities,
the noise quantile, the Wiener gains, the previous noise and magnitude spectra, the LRT average, the pause
spectrum, the seven features, the model parameters re-estimated every 500 frames and the prior speech
probability -- is IDENTICAL to the field of the reference's struct.  (The double-precision library functions
behind log / exp / tanh differ from glibc's in the last place of a double; rounded to float that shows once
in ~2^28 evaluations, a one-ulp ripple that dies out within a few frames.)"""
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
    for s in range(n):
        _, tr = ref_ns_trace(reflib, fs, mode, x[s])
        g = gpu_nsf_state(lib, b._handles[s], fs)
        # the whole decision-directed recursion: trackers, densities, noise quantile, Wiener gains, previous noise
        # and magnitude, LRT average, pause spectrum, the seven features, the model parameters and the prior
        for key in ("lquantile", "density", "magn", "quantile", "smooth", "noisePrev", "logLrt", "magnAvgPause",
                    "featureData", "priorModelPars"):
            assert bits_equal(g[key], tr[key][-1]), "stream %d: %s differs from the reference struct (max |d| %g)" % (
                s, key, np.abs(np.asarray(g[key], np.float64) - tr[key][-1]).max())
        assert g["prior"] == float(tr["prior"][-1])
    b.close()
