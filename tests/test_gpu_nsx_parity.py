"""GPU parity of the fixed-point suppressor (WebRtcNsx_*): int16 output must be bit-identical to
the compiled reference (BASELINE.json config 3: 8 kHz and 16 kHz, every policy)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fs,mode,frames", [(16000, 2, 1300), (8000, 2, 1300), (16000, 0, 260), (16000, 1, 260),
                                            (16000, 3, 260), (8000, 0, 260), (8000, 1, 260), (8000, 3, 260)])
def test_batch_bit_exact_all_stream_classes(nslib, reflib, fs, mode, frames):
    """8 synthetic stream classes (incl. digital silence, clipping, delayed start) through
    WebRtcNsx_ProcessBatch in ragged chunks; 1300 frames cross two 512-frame threshold windows."""
    n = 16
    fl = fs // 100
    x = nslib.synth_pcm_host(n, fs, frames * fl)
    b = nslib.NsBatch(n, fs, mode, fixed=True)
    out = np.zeros_like(x)
    f0 = 0
    for chunk in [1, 3, 17, 100, 10 ** 9]:
        nf = min(chunk, frames - f0)
        if nf <= 0:
            break
        out[:, f0 * fl:(f0 + nf) * fl] = b.process(x[:, f0 * fl:(f0 + nf) * fl])
        f0 += nf
    for s in range(n):
        ref = reflib.nsx(fs, mode, x[s])
        assert np.array_equal(ref, out[s]), "stream %d differs at sample %d" % (s, int(np.nonzero(ref != out[s])[0][0]))
    b.close()


def test_single_stream_api_bit_exact(nslib, reflib):
    fs, mode, frames = 16000, 2, 80
    x = nslib.synth_pcm_host(3, fs, frames * 160)[2]
    ns = nslib.NoiseSuppressorX()
    assert ns.init(fs) == 0 and ns.set_policy(mode) == 0
    out = np.zeros_like(x)
    for f in range(frames):
        out[f * 160:(f + 1) * 160] = ns.process([x[f * 160:(f + 1) * 160]])[0]
    assert np.array_equal(out, reflib.nsx(fs, mode, x))
    ns.free()


def test_error_behaviour(nslib):
    ns = nslib.NoiseSuppressorX()
    assert ns.init(22050) == -1          # nsx_core.c:640-644
    assert ns.init(8000) == 0
    assert ns.set_policy(7) == -1        # nsx_core.c:787-789
    assert ns.set_policy(1) == 0
    ns.free()
    assert nslib.load_library().WebRtcNsx_Init(None, 16000) == -1


def test_large_batch_matches_small_batches(nslib, reflib):
    """2048 streams in one launch: every 64th stream checked against the reference."""
    fs, mode, frames, n = 16000, 2, 120, 2048
    x = nslib.synth_pcm_host(n, fs, frames * 160)
    b = nslib.NsBatch(n, fs, mode, fixed=True)
    out = b.process(x)
    for s in range(0, n, 64):
        assert np.array_equal(out[s], reflib.nsx(fs, mode, x[s]))
    b.close()
