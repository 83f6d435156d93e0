"""WebRtcNs[x]_ProcessBatchAsync: consecutive asynchronous host-pointer calls form one pipeline over the
copy engines; their results must be the blocking call's, bit for bit (same kernels, same chunks of
state), whatever the number of calls in flight, the order of the waits, or the calls mixed in between."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pinned(torch, shape):
    return torch.empty(shape, dtype=torch.int16).pin_memory()


@pytest.mark.parametrize("fixed,fs", [(False, 16000), (True, 16000), (False, 8000)])
def test_async_pipeline_equals_blocking(nslib, fixed, fs):
    import torch
    n, mode, F, calls = 96, 2, 60, 5
    fl = fs // 100
    x = nslib.synth_pcm_host(n, fs, calls * F * fl)
    ref_b = nslib.NsBatch(n, fs, mode, fixed=fixed, devices=[0])
    want = np.concatenate([ref_b.process(np.ascontiguousarray(x[:, c * F * fl:(c + 1) * F * fl])) for c in range(calls)], axis=1)
    ref_b.close()

    b = nslib.NsBatch(n, fs, mode, fixed=fixed, devices=[0])
    ins = [_pinned(torch, (n, F * fl)) for _ in range(calls)]
    outs = [_pinned(torch, (n, F * fl)) for _ in range(calls)]
    for c in range(calls):
        ins[c].copy_(torch.from_numpy(np.ascontiguousarray(x[:, c * F * fl:(c + 1) * F * fl])))
        outs[c].fill_(-1)
    # all five calls in flight at once, waited for out of order
    tickets = [b.process_ptr_async(ins[c].data_ptr(), F * fl, outs[c].data_ptr(), F * fl, F) for c in range(calls)]
    assert all(t > 0 for t in tickets) and len(set(tickets)) == calls
    for c in (3, 0, 4, 1, 2):
        b.wait(tickets[c])
    b.wait(tickets[0])      # waiting twice is a no-op
    b.wait(0)
    got = np.concatenate([o.numpy() for o in outs], axis=1)
    assert np.array_equal(got, want)
    b.close()


def test_async_mixed_with_other_calls(nslib):
    """Any other entry point first waits for the batches in flight: blocking batch, single-stream
    probability getter, re-Init and a differently shaped batch right behind asynchronous calls."""
    import torch
    n, fs, mode, F = 64, 16000, 2, 50
    fl = fs // 100
    x = nslib.synth_pcm_host(n, fs, 4 * F * fl)
    a = nslib.NsBatch(n, fs, mode, devices=[0])
    want = np.concatenate([a.process(np.ascontiguousarray(x[:, c * F * fl:(c + 1) * F * fl])) for c in range(4)], axis=1)
    p_want = a.prior_speech_probability(5)
    a.close()

    b = nslib.NsBatch(n, fs, mode, devices=[0])
    ins = [_pinned(torch, (n, F * fl)) for _ in range(4)]
    outs = [_pinned(torch, (n, F * fl)) for _ in range(4)]
    for c in range(4):
        ins[c].copy_(torch.from_numpy(np.ascontiguousarray(x[:, c * F * fl:(c + 1) * F * fl])))
    b.process_ptr_async(ins[0].data_ptr(), F * fl, outs[0].data_ptr(), F * fl, F)
    b.process_ptr_async(ins[1].data_ptr(), F * fl, outs[1].data_ptr(), F * fl, F)
    # blocking call behind two asynchronous ones (never waited for explicitly)
    b.process_ptr(ins[2].data_ptr(), F * fl, outs[2].data_ptr(), F * fl, F)
    assert np.array_equal(np.concatenate([o.numpy() for o in outs[:3]], axis=1), want[:, :3 * F * fl])
    t = b.process_ptr_async(ins[3].data_ptr(), F * fl, outs[3].data_ptr(), F * fl, F)
    assert b.prior_speech_probability(5) == p_want      # drains the pipeline first
    assert np.array_equal(outs[3].numpy(), want[:, 3 * F * fl:])
    b.wait(t)
    # a batch of another shape right behind an asynchronous call: the staging is re-cut safely
    b.reset(mode)
    t = b.process_ptr_async(ins[0].data_ptr(), F * fl, outs[0].data_ptr(), F * fl, F)
    half = nslib.NsBatch(n // 2, fs, mode, devices=[0])
    o = half.process(np.ascontiguousarray(x[: n // 2, : 2 * F * fl]))
    assert np.array_equal(o, want[: n // 2, : 2 * F * fl])
    assert np.array_equal(outs[0].numpy(), want[:, : F * fl])
    half.close()
    b.close()


def test_async_48k_blocks_and_returns_ticket_zero(nslib):
    import torch
    n, fs, mode, F = 8, 48000, 2, 20
    fl = fs // 100
    x = nslib.synth_pcm_host(n, fs, F * fl)
    a = nslib.NsBatch(n, fs, mode, fixed=True, devices=[0])
    want = a.process(x)
    a.close()
    b = nslib.NsBatch(n, fs, mode, fixed=True, devices=[0])
    i, o = _pinned(torch, (n, F * fl)), _pinned(torch, (n, F * fl))
    i.copy_(torch.from_numpy(x))
    assert b.process_ptr_async(i.data_ptr(), F * fl, o.data_ptr(), F * fl, F) == 0
    assert np.array_equal(o.numpy(), want)
    b.close()
