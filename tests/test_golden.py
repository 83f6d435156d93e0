"""Committed golden vectors (tests/golden/*.npz, produced by the compiled reference with
tests/golden/make_golden.py): the CPU oracle must reproduce them here; the GPU library must
reproduce them on the box (test_gpu_golden.py) even when oracle/_ref is absent."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, _ptr, judge_float, summarize_parity

FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ns_fs*.npz")))


def load(path, synth):
    g = np.load(path)
    fs, mode = int(g["fs"]), int(g["mode"])
    frames = g["nsx_out"].shape[1] // (fs // 100)
    x = np.stack([synth(1, fs, frames * (fs // 100), base_seed=int(g["seed"]), first_stream=int(s))[0]
                  for s in g["streams"]])
    # the generator must still produce the input the vectors were made from
    assert np.array_equal(x.astype(np.int64).sum(1), g["pcm_in_sum"])
    assert np.array_equal((x.astype(np.int64) ** 2).sum(1), g["pcm_in_sqsum"])
    return g, fs, mode, frames, x


@pytest.fixture(scope="module")
def oracle():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(ROOT, "oracle", "liboracle_ns.so"))


def test_golden_files_present():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_oracle_reproduces_golden(path, oracle, nslib_host_synth):
    g, fs, mode, frames, x = load(path, nslib_host_synth)
    for i in range(x.shape[0]):
        xs = np.ascontiguousarray(x[i])
        out = np.zeros_like(xs)
        assert oracle.nsx_oracle_run(fs, mode, frames, _ptr(xs), _ptr(out)) == 0
        assert np.array_equal(out, g["nsx_out"][i]), "NSx golden mismatch, stream %d" % i
        of = np.zeros(len(xs), np.float32)
        pp = np.zeros(frames, np.float32)
        assert oracle.nsf_oracle_run(fs, mode, frames, _ptr(xs), _ptr(of), _ptr(pp)) == 0
        assert np.array_equal(g["ns_out"][i].view(np.uint32), of.view(np.uint32)), "float NS golden mismatch, stream %d" % i
        assert np.array_equal(np.asarray(g["ns_prior_prob"][i], np.float32).view(np.uint32), pp.view(np.uint32))
