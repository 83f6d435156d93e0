"""Python mirror of the reference's NS interface (same names, argument meaning and error
behaviour as ns/include/noise_suppression.h and noise_suppression_x.h) plus the batched entry
point.  All compute happens in libwebrtc_ns_b200.so on the GPU."""
import ctypes as C

import numpy as np

from .capi import load_library


class NsError(RuntimeError):
    pass


def frame_len(fs):
    return fs // 100


def num_bands(fs):
    return {8000: 1, 16000: 1, 32000: 2, 48000: 3}[fs]


def _err(lib, what):
    return NsError("%s failed: %s" % (what, lib.WebRtcNsB200_LastError().decode()))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class _Single:
    _prefix = None

    def __init__(self, device=None):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.fs = 0
        if device is not None:
            self._lib.WebRtcNsB200_SetCreateDevice(int(device))
        rc = getattr(self._lib, self._prefix + "_Create")(C.byref(self._h))
        if device is not None:
            self._lib.WebRtcNsB200_SetCreateDevice(-1)
        if rc != 0:
            raise _err(self._lib, self._prefix + "_Create")

    def init(self, fs):
        """WebRtcNs_Init / WebRtcNsx_Init: returns 0 or -1 like the reference."""
        rc = getattr(self._lib, self._prefix + "_Init")(self._h, int(fs) & 0xFFFFFFFF)
        if rc == 0:
            self.fs = fs
        return rc

    def set_policy(self, mode):
        return getattr(self._lib, self._prefix + "_set_policy")(self._h, int(mode))

    def free(self):
        if self._h:
            getattr(self._lib, self._prefix + "_Free")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class NoiseSuppressor(_Single):
    """One float-NS handle (WebRtcNs_*)."""
    _prefix = "WebRtcNs"

    def analyze(self, frame):
        f = np.ascontiguousarray(frame, dtype=np.float32)
        self._lib.WebRtcNs_Analyze(self._h, _ptr(f))

    def process(self, bands):
        """bands: list of num_bands float32 frames (int16 scale). Returns the output frames."""
        ins = [np.ascontiguousarray(b, dtype=np.float32) for b in bands]
        outs = [np.zeros_like(b) for b in ins]
        n = len(ins)
        pi = (C.c_void_p * n)(*[b.ctypes.data for b in ins])
        po = (C.c_void_p * n)(*[b.ctypes.data for b in outs])
        self._lib.WebRtcNs_Process(self._h, pi, n, po)
        return outs

    def prior_speech_probability(self):
        return float(self._lib.WebRtcNs_prior_speech_probability(self._h))


class NoiseSuppressorX(_Single):
    """One fixed-point NSx handle (WebRtcNsx_*)."""
    _prefix = "WebRtcNsx"

    def process(self, bands):
        ins = [np.ascontiguousarray(b, dtype=np.int16) for b in bands]
        outs = [np.zeros_like(b) for b in ins]
        n = len(ins)
        pi = (C.c_void_p * n)(*[b.ctypes.data for b in ins])
        po = (C.c_void_p * n)(*[b.ctypes.data for b in outs])
        self._lib.WebRtcNsx_Process(self._h, pi, n, po)
        return outs


class NsBatch:
    """N independent streams behind WebRtcNs[x]_ProcessBatch.

    devices: list of GPU indices; streams are sharded in contiguous slices
    (stream s -> devices[s // ceil(N/G)])."""

    def __init__(self, n_streams, fs, mode, fixed=False, devices=None):
        self._lib = load_library()
        self.n = int(n_streams)
        self.fs = int(fs)
        self.fixed = bool(fixed)
        self._p = "WebRtcNsx" if fixed else "WebRtcNs"
        self._handles = (C.c_void_p * self.n)()
        create = getattr(self._lib, self._p + "_Create")
        per = None
        if devices:
            per = -(-self.n // len(devices))
        for s in range(self.n):
            if devices:
                self._lib.WebRtcNsB200_SetCreateDevice(int(devices[s // per]))
            h = C.c_void_p()
            if create(C.byref(h)) != 0:
                raise _err(self._lib, self._p + "_Create")
            self._handles[s] = h
        if devices:
            self._lib.WebRtcNsB200_SetCreateDevice(-1)
        if getattr(self._lib, self._p + "_InitBatch")(self._handles, self.n, self.fs & 0xFFFFFFFF, int(mode)) != 0:
            raise _err(self._lib, self._p + "_InitBatch")

    def reset(self, mode):
        if getattr(self._lib, self._p + "_InitBatch")(self._handles, self.n, self.fs & 0xFFFFFFFF, int(mode)) != 0:
            raise _err(self._lib, self._p + "_InitBatch")

    def process(self, pcm_in, pcm_out=None):
        """pcm_in: int16 [n_streams, frames * fs/100] host array (numpy). Returns int16 output."""
        x = np.ascontiguousarray(pcm_in, dtype=np.int16)
        assert x.ndim == 2 and x.shape[0] == self.n and x.shape[1] % frame_len(self.fs) == 0
        out = np.empty_like(x) if pcm_out is None else pcm_out
        frames = x.shape[1] // frame_len(self.fs)
        rc = getattr(self._lib, self._p + "_ProcessBatch")(
            self._handles, self.n, _ptr(x), x.shape[1], _ptr(out), out.shape[1], frames)
        if rc != 0:
            raise _err(self._lib, self._p + "_ProcessBatch")
        return out

    def process_ptr(self, in_ptr, in_stride, out_ptr, out_stride, frames):
        """Host pointers (e.g. pinned torch tensors): the raw C call."""
        rc = getattr(self._lib, self._p + "_ProcessBatch")(
            self._handles, self.n, C.c_void_p(in_ptr), in_stride, C.c_void_p(out_ptr), out_stride, frames)
        if rc != 0:
            raise _err(self._lib, self._p + "_ProcessBatch")

    def process_ptr_async(self, in_ptr, in_stride, out_ptr, out_stride, frames):
        """Page-locked host pointers; returns a ticket at once (WebRtcNs[x]_ProcessBatchAsync).  The
        buffers belong to the library until wait(ticket)."""
        t = C.c_uint64(0)
        rc = getattr(self._lib, self._p + "_ProcessBatchAsync")(
            self._handles, self.n, C.c_void_p(in_ptr), in_stride, C.c_void_p(out_ptr), out_stride, frames, C.byref(t))
        if rc != 0:
            raise _err(self._lib, self._p + "_ProcessBatchAsync")
        return t.value

    def wait(self, ticket):
        if self._lib.WebRtcNsB200_WaitBatch(C.c_uint64(ticket)) != 0:
            raise _err(self._lib, "WebRtcNsB200_WaitBatch")

    def process_device(self, in_ptr, in_stride, out_ptr, out_stride, frames, stream=0):
        """Device pointers; enqueues on `stream` (cudaStream_t as int) and returns immediately."""
        rc = getattr(self._lib, self._p + "_ProcessBatchDevice")(
            self._handles, self.n, C.c_void_p(in_ptr), in_stride, C.c_void_p(out_ptr), out_stride,
            frames, C.c_void_p(stream))
        if rc != 0:
            raise _err(self._lib, self._p + "_ProcessBatchDevice")

    def process_bands_f32(self, bands_in):
        """float32 [n_streams, frames, num_bands, frame_len]: batch mirror of WebRtcNs_Process."""
        assert not self.fixed
        x = np.ascontiguousarray(bands_in, dtype=np.float32)
        n, frames, nb, fl = x.shape
        out = np.empty_like(x)
        rc = self._lib.WebRtcNs_ProcessBatchBandsF32(self._handles, n, nb, _ptr(x), frames * nb * fl,
                                                     _ptr(out), frames * nb * fl, frames)
        if rc != 0:
            raise _err(self._lib, "WebRtcNs_ProcessBatchBandsF32")
        return out

    def prior_speech_probability(self, s):
        assert not self.fixed
        return float(self._lib.WebRtcNs_prior_speech_probability(self._handles[s]))

    def close(self):
        free = getattr(self._lib, self._p + "_Free")
        for s in range(self.n):
            if self._handles[s]:
                free(self._handles[s])
                self._handles[s] = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_pcm_host(n_streams, fs, n_samples, base_seed=1234, first_stream=0, first_sample=0):
    """Deterministic synthetic PCM (csrc/pcm_synth.h) as int16 [n_streams, n_samples]."""
    lib = load_library()
    out = np.empty((n_streams, n_samples), np.int16)
    for s in range(n_streams):
        lib.WebRtcNsB200_SynthPcmHost(_ptr(out[s]), first_stream + s, fs, first_sample, n_samples, base_seed)
    return out


def run_generated_job(batch, frames, frames_per_launch, first_stream=0, base_seed=1234, sample=(), stream=None,
                      on_launch=None):
    """A job too large to hold as PCM (BASELINE config 5: 65 536 streams x 10 min): per
    chunk of `frames_per_launch` frames the synthetic PCM of every stream of `batch` is generated on the
    device, processed, and reduced to a running per-stream (sum, energy) checksum; input and output of
    the streams listed in `sample` are also kept (host arrays) for comparison with the reference.
    Everything is enqueued on one CUDA stream.  Returns (sums int64 [n, 2] numpy, {s: (in, out)}).
    on_launch(i, ev_before, ev_after): optional hook receiving CUDA events around each process call."""
    import torch
    lib = batch._lib
    n, fs = batch.n, batch.fs
    fl = frame_len(fs)
    F = int(frames_per_launch)
    st = stream or torch.cuda.Stream()
    bufs = [(torch.empty((n, F * fl), dtype=torch.int16, device="cuda"),
             torch.empty((n, F * fl), dtype=torch.int16, device="cuda")) for _ in range(2)]
    sums = torch.zeros((n, 2), dtype=torch.int64, device="cuda")
    sample = list(sample)
    idx = torch.tensor(sample, dtype=torch.long, device="cuda") if sample else None
    keep_in = torch.empty((len(sample), frames * fl), dtype=torch.int16, device="cuda") if sample else None
    keep_out = torch.empty_like(keep_in) if sample else None
    with torch.cuda.stream(st):
        i = 0
        for f0 in range(0, frames, F):
            nf = min(F, frames - f0)
            x, y = bufs[i % 2]
            rc = lib.WebRtcNsB200_SynthPcmDevice(C.c_void_p(x.data_ptr()), F * fl, n, first_stream, fs, f0 * fl,
                                                 nf * fl, base_seed, C.c_void_p(st.cuda_stream))
            if rc != 0:
                raise _err(lib, "WebRtcNsB200_SynthPcmDevice")
            ev = None
            if on_launch is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record(st)
            batch.process_device(x.data_ptr(), F * fl, y.data_ptr(), F * fl, nf, st.cuda_stream)
            if ev is not None:
                ev[1].record(st)
                on_launch(i, ev[0], ev[1])
            rc = lib.WebRtcNsB200_ChecksumAccumulateDevice(C.c_void_p(y.data_ptr()), F * fl, n, nf * fl,
                                                           C.c_void_p(sums.data_ptr()), C.c_void_p(st.cuda_stream))
            if rc != 0:
                raise _err(lib, "WebRtcNsB200_ChecksumAccumulateDevice")
            if sample:
                keep_in[:, f0 * fl:(f0 + nf) * fl] = x[idx, :nf * fl]
                keep_out[:, f0 * fl:(f0 + nf) * fl] = y[idx, :nf * fl]
            i += 1
    st.synchronize()
    out = {}
    if sample:
        ki, ko = keep_in.cpu().numpy(), keep_out.cpu().numpy()
        out = {s: (ki[j], ko[j]) for j, s in enumerate(sample)}
    return sums.cpu().numpy(), out
