"""CPU-only calibration of what float parity can mean: the reference compared with ITSELF.

oracle/_ref/libns_ref.so is the reference built with every multiply/add rounded separately
(-ffp-contract=off); libns_ref_fma.so is the same unmodified source with FMA contraction allowed in
ns_core.c and fft4g.c (gcc's default under -march=native).  Both are conforming builds.  On our
synthetic streams they agree to ~0.05 LSB until a knife-edge comparison goes the other way, after
which they differ by tens of LSB for seconds.  The GPU parity gates in conftest.py are set from this."""
import numpy as np

from conftest import judge_float


def test_reference_self_consistency(nslib_host_synth, reflib, reflib_fma):
    fs, mode, n, frames = 16000, 2, 16, 1200
    x = nslib_host_synth(n, fs, frames * 160)
    res = []
    for s in range(n):
        a = reflib.ns(fs, mode, x[s])[0]
        b = reflib_fma.ns(fs, mode, x[s])[0]
        res.append(judge_float(a, b))
    n_strict = sum(1 for r in res if r[0])
    worst = max(r[2] for r in res)
    print("reference plain vs FMA-contracted: %d/%d streams within 1e-4 FS & 90 dB, worst %.1f LSB, %.1f dB"
          % (n_strict, n, worst, min(r[3] for r in res)))
    # the point of the test: the strict tolerance does NOT hold between two builds of the reference
    assert n_strict < n
    assert all(r[1] for r in res)      # ... but the envelope does
    assert n_strict >= 0.6 * n
