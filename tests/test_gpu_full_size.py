"""BASELINE.json configurations at their full stream counts and durations, PCM generated and kept
on the device: a sampled subset of the streams is read back and checked against the
compiled reference (bit-exact for the fixed-point suppressor and the band split, float gates for
the float suppressor), and size-independent properties cover every stream: the per-stream checksum
of the whole batch does not depend on how the frames are cut into launches."""
import ctypes as C

import numpy as np
import pytest

from conftest import judge_float, summarize_parity

pytestmark = pytest.mark.gpu


def _run(nslib, torch, n, fs, mode, fixed, frames, per_launch, sample):
    lib = nslib.load_library()
    fl = fs // 100
    ns = frames * fl
    st = torch.cuda.Stream()
    x = torch.empty((n, ns), dtype=torch.int16, device="cuda")
    y = torch.empty_like(x)
    assert lib.WebRtcNsB200_SynthPcmDevice(C.c_void_p(x.data_ptr()), ns, n, 0, fs, 0, ns, 1234,
                                           C.c_void_p(st.cuda_stream)) == 0
    b = nslib.NsBatch(n, fs, mode, fixed=fixed, devices=[0])
    for f0 in range(0, frames, per_launch):
        nf = min(per_launch, frames - f0)
        b.process_device(x.data_ptr() + 2 * f0 * fl, ns, y.data_ptr() + 2 * f0 * fl, ns, nf, st.cuda_stream)
    sums = torch.zeros((n, 2), dtype=torch.int64, device="cuda")
    assert lib.WebRtcNsB200_ChecksumDevice(C.c_void_p(y.data_ptr()), ns, n, ns, C.c_void_p(sums.data_ptr()),
                                           C.c_void_p(st.cuda_stream)) == 0
    st.synchronize()
    torch.cuda.synchronize()
    out = {s: (x[s].cpu().numpy(), y[s].cpu().numpy()) for s in sample}
    b.close()
    return sums.cpu().numpy(), out


@pytest.mark.timeout(900)
def test_config2_float_4096_streams_60s(nslib, reflib):
    """configs[1]: 4096 x 16 kHz x 60 s, float NS policy 2 (the bench workload)."""
    import torch
    n, fs, mode, frames = 4096, 16000, 2, 6000
    sample = [0, 1, 2, 3, 4, 5, 6, 7, 1031, 2050, 3333, 4095]
    sums_a, out = _run(nslib, torch, n, fs, mode, False, frames, 100, sample)
    res = []
    for s in sample:
        xin, yout = out[s]
        assert np.array_equal(xin, nslib.synth_pcm_host(1, fs, frames * (fs // 100), first_stream=s)[0])
        _, refi, _ = reflib.ns(fs, mode, xin)
        res.append(judge_float(refi.astype(np.float32), yout.astype(np.float32), slack=1.0))
    summarize_parity(res, "config 2 full size", 1.0, max_abs=1.0)
    # launch partition independence over ALL streams: 60 launches of 100 frames == 8 launches of 750+
    sums_b, _ = _run(nslib, torch, n, fs, mode, False, frames, 777, [])
    assert np.array_equal(sums_a, sums_b)
    assert len(np.unique(sums_a[:, 1])) > n // 2      # the streams really are different signals


@pytest.mark.timeout(900)
@pytest.mark.parametrize("fs", [8000, 16000])
def test_config3_fixed_8192_streams_60s(nslib, reflib, fs):
    """configs[2]: 8192 x NSx policy 2 x 60 s at 8 and 16 kHz, memcmp against the reference."""
    import torch
    n, mode, frames = 8192, 2, 6000
    sample = [0, 3, 4, 5, 6, 4097, 8191]
    sums_a, out = _run(nslib, torch, n, fs, mode, True, frames, 100, sample)
    for s in sample:
        xin, yout = out[s]
        assert np.array_equal(reflib.nsx(fs, mode, xin), yout), "stream %d" % s
    sums_b, _ = _run(nslib, torch, n, fs, mode, True, frames, 1234, [])
    assert np.array_equal(sums_a, sums_b)


@pytest.mark.timeout(900)
def test_config4_48k_2048_streams(nslib, reflib):
    """configs[3]: 2048 x 48 kHz x 3 bands through resampler + QMF split/merge, 30 s: fixed-point
    chain bit-exact on the sample, float chain inside the float gates."""
    import torch
    n, fs, mode, frames = 2048, 48000, 2, 3000
    sample = [0, 2, 5, 1029, 2047]
    sums_a, out = _run(nslib, torch, n, fs, mode, True, frames, 50, sample)
    for s in sample:
        xin, yout = out[s]
        assert np.array_equal(reflib.nsx(fs, mode, xin), yout), "stream %d" % s
    sums_b, _ = _run(nslib, torch, n, fs, mode, True, frames, 333, [])
    assert np.array_equal(sums_a, sums_b)
    _, outf = _run(nslib, torch, n, fs, mode, False, frames, 50, sample)
    res = []
    for s in sample:
        xin, yout = outf[s]
        _, refi, _ = reflib.ns(fs, mode, xin)
        res.append(judge_float(refi.astype(np.float32), yout.astype(np.float32), slack=1.0))
    summarize_parity(res, "config 4 full size", 1.0)


@pytest.mark.timeout(1200)
def test_config5_shard_8192_streams_600s(nslib, reflib):
    """configs[4]: 65 536 streams x 10 min at 16 kHz, sharded 8192 per GPU on 8 B200 -- one GPU's shard
    at full length (8192 x 60 000 frames = 4.9e8 stream-frames).  1.26 TB of PCM each way in the whole
    job, so the PCM is generated on the device chunk by chunk and the output reduced to a running
    per-stream (sum, energy) checksum; sampled streams are kept and checked against the
    compiled reference over the whole 10 minutes, and the checksums of ALL streams must not depend on
    how the 60 000 frames are cut into launches (state carried through 600 launches == through 80)."""
    n, fs, mode, frames = 8192, 16000, 2, 60000
    first = 5 * 8192                       # the shard of rank 5 of 8: global stream indices
    sample = [0, 3, 4102, 8191]
    b = nslib.NsBatch(n, fs, mode, devices=[0])
    sums_a, out = nslib.run_generated_job(b, frames, 100, first_stream=first, sample=sample)
    res = []
    for s in sample:
        xin, yout = out[s]
        assert np.array_equal(xin, nslib.synth_pcm_host(1, fs, frames * (fs // 100), first_stream=first + s)[0])
        _, refi, _ = reflib.ns(fs, mode, xin)
        res.append(judge_float(refi.astype(np.float32), yout.astype(np.float32), slack=1.0))
        # the kept output is what the checksum saw
        y64 = yout.astype(np.int64)
        assert sums_a[s, 0] == y64.sum() and sums_a[s, 1] == (y64 * y64).sum()
    summarize_parity(res, "config 5 shard, 10 min", 1.0, max_abs=1.0)
    b.reset(mode)
    sums_b, _ = nslib.run_generated_job(b, frames, 750, first_stream=first)
    b.close()
    assert np.array_equal(sums_a, sums_b)
    assert len(np.unique(sums_a[:, 1])) > n // 2
