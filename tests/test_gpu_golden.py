"""GPU library against the committed golden vectors (no oracle/_ref needed on the box)."""
import os

import numpy as np
import pytest

from conftest import FLOAT_MAX_ABS, FLOAT_MIN_SNR, judge_float, summarize_parity
from test_golden import FILES, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_gpu_reproduces_golden(path, nslib):
    g, fs, mode, frames, x = load(path, nslib.synth_pcm_host)
    n, fl = x.shape[0], fs // 100
    bx = nslib.NsBatch(n, fs, mode, fixed=True)
    assert np.array_equal(bx.process(x), g["nsx_out"]), "NSx int16 output differs from the reference's"
    bx.close()
    bf = nslib.NsBatch(n, fs, mode)
    out = bf.process_bands_f32(x.astype(np.float32).reshape(n, frames, 1, fl)).reshape(n, -1)
    res = [judge_float(g["ns_out"][i], out[i]) for i in range(n)]
    for i in range(n):
        if res[i][0]:
            assert abs(bf.prior_speech_probability(i) - float(g["ns_prior_prob"][i][-1])) <= 5e-4
    summarize_parity(res, "float GPU vs golden %s" % os.path.basename(path), 1.0, max_abs=FLOAT_MAX_ABS, min_snr=FLOAT_MIN_SNR)
    bf.close()
