"""examples/test_ns_module (Linux clone of the reference CLI driver) end to end on a WAV file:
BASELINE.json config 1 (single 16 kHz mono WAV through float NS on 10 ms frames)."""
import os
import subprocess
import wave

import numpy as np
import pytest

from conftest import ROOT, judge_float

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "audiosignalprocess_b200", "test_ns_module")


def _write_wav(path, x, fs, channels):
    with wave.open(path, "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(2)
        w.setframerate(fs)
        w.writeframes(x.astype("<i2").tobytes())


def _read_wav(path):
    with wave.open(path, "rb") as w:
        return np.frombuffer(w.readframes(w.getnframes()), "<i2").copy(), w.getframerate(), w.getnchannels()


@pytest.mark.parametrize("fs,fixed,block", [(16000, False, 1), (16000, True, 1), (48000, False, 25), (8000, True, 50)])
def test_cli_matches_reference(tmp_path, nslib, reflib, fs, fixed, block):
    from audiosignalprocess_b200 import build
    build.build_examples()
    fl, frames = fs // 100, 150
    x = nslib.synth_pcm_host(1, fs, frames * fl + 37, first_stream=2)[0]      # ragged tail of 37 samples
    src, dst = str(tmp_path / "in.wav"), str(tmp_path / "out.wav")
    _write_wav(src, x, fs, 1)
    cmd = [EXE, src, dst, "--block", str(block)] + (["--fixed"] if fixed else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    y, rate, ch = _read_wav(dst)
    assert rate == fs and ch == 1 and len(y) == len(x)
    body = x[:frames * fl]
    if fixed:
        assert np.array_equal(y[:frames * fl], reflib.nsx(fs, 1, body))     # the driver's policy is kModerate = 1
    else:
        _, refi, _ = reflib.ns(fs, 1, body)
        assert judge_float(refi, y[:frames * fl], slack=1.0)[0]
    assert np.array_equal(y[frames * fl:], x[frames * fl:])                   # partial frame passes through
