#!/usr/bin/env python
"""Headline benchmark: NS audio-seconds per second over N streams (BASELINE.json metric).

Workload (configs[1]): 4096 independent 16 kHz mono streams per GPU, float NS policy 2,
synthetic PCM (csrc/pcm_synth.h, 8 stream classes).  One step = one WebRtcNs_ProcessBatch
call walking F = --frames-per-step 10-ms frames of every stream; with the default
--steps 60 and F = 100 the timed region is the configuration's full 60 s of audio.

  value     PCM resident in HBM, WebRtcNs_ProcessBatchDevice, CUDA events on the launch stream
  e2e       same steps through WebRtcNs_ProcessBatch with pinned HOST buffers
            (H2D + kernels + D2H inside the timed region)
  roofline  HBM: algorithmic bytes B(F) = IO + 2*S_hot/F per stream-frame; at F = 100 the
            kernel is issue-bound (roofline.issue), so the HBM fraction is small by construction
  roofline_tick  the same kernel at F = 1 over --tick-streams streams (the API-faithful 10 ms tick, every
            launch moves the whole per-stream state): the regime the HBM roofline bounds
  cpu_baseline / --impl reference: the unmodified reference C (oracle/_ref) on the host cores, driven like the
            GPU arm (persistent handles, same warm-up, same frames of every stream; never loads the product .so)

N > 1: one process per GPU under torchrun, every rank owns its own 4096 streams (weak scaling,
no data-path collective -- streams are independent); barrier + max over ranks for the time.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S_HOT = {("float", 16000): 7120, ("float", 8000): 3664, ("float", 48000): 8448, ("float", 32000): 7888,
         ("fixed", 16000): 4414, ("fixed", 8000): 2302, ("fixed", 32000): 5182, ("fixed", 48000): 5742}


def algorithmic_bytes(kind, fs, frames_per_launch):
    io = 2 * 2 * (fs // 100)            # int16 in + int16 out per stream-frame
    return io + 2.0 * S_HOT[(kind, fs)] / frames_per_launch


def _profile_summary(kernel, frames_per_launch, streams):
    """The committed ncu summary (profiles/r<round>_*.json, newest round first) of the SAME kernel and
    configuration, with its file name; (None, None) if there is none."""
    tag = {"nsf_process_kernel": "nsf", "nsx_process_kernel": "nsx"}.get(kernel)
    for rnd in ("r2", "r1"):
        for name in ("%s_%s_kernel_F%d.json" % (rnd, tag, frames_per_launch),
                     "%s_%s_kernel_F%d_%d.json" % (rnd, tag, frames_per_launch, streams)):
            try:
                d = json.load(open(os.path.join(ROOT, "profiles", name)))
                if d["units_per_launch"] == streams * frames_per_launch:
                    return d, "profiles/" + name
            except Exception:
                pass
    return None, None


def measured_traffic(kernel, frames_per_launch, streams):
    """DRAM bytes per launch of the dominant kernel from the committed ncu summary of the same configuration."""
    d, _ = _profile_summary(kernel, frames_per_launch, streams)
    return d["dram_bytes_per_launch"] if d else None


def measured_instructions(kernel, frames_per_launch, streams):
    """Warp instructions per stream-frame of the dominant kernel from the committed ncu summary of the
    same configuration (for the issue-rate view of a kernel that is not HBM-bound)."""
    d, _ = _profile_summary(kernel, frames_per_launch, streams)
    return d["warp_instructions_per_unit"] if d else None


def pcie_probe(torch, dev, seconds=1.0, mbytes=128):
    """Host<->device copy ceiling of THIS rank's GPU as the end-to-end path uses it: pinned buffers, an H2D
    and a D2H copy in flight together on two streams, SUSTAINED for `seconds` (the caller puts a barrier in
    front, so under torchrun every rank copies during the same second: the ranks share the host's memory /
    PCIe fabric, and a best-of-N of short bursts measures the moments the others were idle).  Returns the
    mean GB/s per direction."""
    n = mbytes * 1000 * 1000 // 2
    h_in = torch.empty(n, dtype=torch.int16).pin_memory()
    h_out = torch.empty(n, dtype=torch.int16).pin_memory()
    d_a = torch.empty(n, dtype=torch.int16, device=dev)
    d_b = torch.empty(n, dtype=torch.int16, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    reps = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        reps += 1
    return reps * n * 2 / (time.perf_counter() - t0) / 1e9


def extra_leg(pkg, lib, torch, dev, stream, kind, fs, streams, F, mode, steps=10, warmup=3):
    """One more configuration of BASELINE.json's list on this GPU, device-resident like `value`: timed with
    CUDA events over `steps` launches, its HBM roofline, and a parity flag from sampled streams checked
    against the compiled reference (oracle/_ref) on the first F frames."""
    import numpy as np
    fl = fs // 100
    fixed = kind == "fixed"
    total = (warmup + steps) * F * fl
    pcm_in = torch.empty((streams, total), dtype=torch.int16, device=dev)
    pcm_out = torch.empty_like(pcm_in)
    rc = lib.WebRtcNsB200_SynthPcmDevice(C.c_void_p(pcm_in.data_ptr()), total, streams, 0, fs, 0, total, 1234,
                                         C.c_void_p(stream.cuda_stream))
    assert rc == 0, lib.WebRtcNsB200_LastError()
    batch = pkg.NsBatch(streams, fs, mode, fixed=fixed, devices=[dev.index])

    def step(i):
        off = i * F * fl * 2
        batch.process_device(pcm_in.data_ptr() + off, total, pcm_out.data_ptr() + off, total, F, stream.cuda_stream)

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        step(warmup + i)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # parity: sampled streams, the first F frames (fresh state), against the unmodified reference
    parity = "unchecked (oracle/_ref missing)"
    ref = ref_lib()
    if ref is not None:
        sample = sorted({0, 1, 2, 3, streams // 2, streams - 1})
        idx = torch.tensor(sample, dtype=torch.long, device=dev)
        xin = pcm_in[idx, :F * fl].cpu().numpy()
        got = pcm_out[idx, :F * fl].cpu().numpy()
        worst = 0
        for k in range(len(sample)):
            want = np.zeros(F * fl, np.int16)
            x1 = np.ascontiguousarray(xin[k])
            if fixed:
                ref.ref_nsx_run(fs, mode, F, x1.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
            else:
                ref.ref_ns_run(fs, mode, F, x1.ctypes.data_as(C.c_void_p), None, want.ctypes.data_as(C.c_void_p), None)
            worst = max(worst, int(np.abs(got[k].astype(np.int32) - want.astype(np.int32)).max()))
        tol = 0 if fixed else (1 if fs <= 16000 else 4)
        parity = ("bit-exact" if worst == 0 else "max |diff| %d LSB" % worst) + " on %d sampled streams x %d frames vs the reference%s" % (
            len(sample), F, "" if worst <= tol else " -- OUT OF TOLERANCE")
    batch.close()
    del pcm_in, pcm_out
    torch.cuda.empty_cache()
    peak, which = measured_peak_gbs()
    bsf = algorithmic_bytes(kind, fs, F)
    achieved = bsf * streams * F / (ms * 1e-3) / 1e9
    return {"workload": "%d streams x %d Hz, %s policy %d, F=%d frames per launch" % (
                streams, fs, "WebRtcNsx fixed" if fixed else "WebRtcNs float", mode, F),
            "value": streams * F * 0.01 / (ms * 1e-3), "unit": "audio-s/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "bytes_per_stream_frame": bsf, "peak_source": which},
            "parity": parity}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed regions (B200_PROFILING.md's clocks line), read
    in-process through NVML: spawning nvidia-smi ten times a second stalls the CUDA driver calls of
    the end-to-end pipeline (measured: 5.9 ms instead of 3.6 ms per step).  nvidia-smi is the
    fallback when NVML cannot be loaded."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []      # (sm_mhz, sm_max_mhz, [reason flags])
        self.stop_flag = False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
        except Exception:
            self.nv = None

    def sample_nvml(self):
        nv = self.nv
        sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(
            nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
        self.samples.append((int(sm), int(mx), [bool(r & b) for b in bits]))

    def sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                            "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        f = [t.strip() for t in o.strip().split(",")]
        if len(f) >= 6 and f[0].isdigit() and f[1].isdigit():
            self.samples.append((int(f[0]), int(f[1]), [t.lower().startswith("active") for t in f[2:6]]))

    def run(self):
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    self.sample_nvml()
                else:
                    self.sample_smi()
            except Exception:
                pass
            time.sleep(0.05 if self.nv is not None else 0.5)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2][i] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(s[1] for s in self.samples),
                "reasons": reasons, "samples": len(self.samples), "via": "nvml" if self.nv is not None else "nvidia-smi"}


def ref_lib():
    path = os.path.join(ROOT, "oracle", "_ref", "libns_ref.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_batch_create.restype = C.c_void_p
    lib.ref_batch_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_batch_step.restype = C.c_double
    lib.ref_batch_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    lib.ref_batch_free.argtypes = [C.c_void_p]
    lib.ref_synth_pcm.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.c_uint32, C.c_int]
    return lib


def cpu_reference_steps(kind, fs, mode, streams, frames_per_step, steps, warmup, threads, first_stream=0,
                        time_limit=150.0):
    """The unmodified reference (oracle/_ref) on `threads` host threads, driven like the GPU arm: persistent
    handles (one per stream, created and initialised once), `warmup` untimed steps, then `steps` timed steps
    of `frames_per_step` frames over all `streams` streams -- the same frames of the same streams the GPU
    arm's end-to-end leg times (it alternates the PCM of the first two steps; so does this).  Streams are
    handed to the threads one at a time from a shared counter.  PCM comes from csrc/pcm_synth.h compiled
    into the shim: the product library is never loaded.  Returns (audio-s/s, steps actually timed)."""
    import numpy as np
    lib = ref_lib()
    if lib is None:
        return None, 0
    fl = fs // 100
    n1 = frames_per_step * fl
    x = [np.zeros((streams, n1), np.int16) for _ in range(2)]
    for k in range(2):
        lib.ref_synth_pcm(x[k].ctypes.data, n1, streams, first_stream, fs, k * n1, n1, 1234, threads)
    out = np.zeros_like(x[0])
    b = lib.ref_batch_create(1 if kind == "fixed" else 0, fs, mode, streams, threads)
    if not b:
        return None, 0
    for i in range(warmup):
        lib.ref_batch_step(b, frames_per_step, x[i & 1].ctypes.data, n1, out.ctypes.data, n1)
    secs = []
    t0 = time.time()
    for i in range(steps):
        secs.append(lib.ref_batch_step(b, frames_per_step, x[(warmup + i) & 1].ctypes.data, n1, out.ctypes.data, n1))
        if time.time() - t0 > time_limit:
            break
    lib.ref_batch_free(b)
    return streams * frames_per_step * 0.01 * len(secs) / sum(secs), len(secs)


def tick_roofline(a, pkg, lib, torch, dev, stream, kind):
    """The API-faithful 10 ms tick (the regime the HBM roofline is meaningful for): ONE frame per
    stream per call over --tick-streams streams, so that every launch loads and stores the whole per-stream
    state (B(1) = IO + 2*S_hot bytes per stream-frame) and the state of all streams (360 MB float / 16 kHz at
    32768 streams) cannot stay in the 126 MB L2.  Timed in steady state, past the 200 start-up frames."""
    n, fl = a.tick_streams, a.fs // 100
    warm, steps = 260, 200   # past the start-up regimes: 50 frames (parametric noise), 200 frames (tracker latch every frame)
    total = (warm + steps) * fl
    pcm_in = torch.empty((n, total), dtype=torch.int16, device=dev)
    pcm_out = torch.empty_like(pcm_in)
    rc = lib.WebRtcNsB200_SynthPcmDevice(C.c_void_p(pcm_in.data_ptr()), total, n, 0, a.fs, 0, total, 4321,
                                         C.c_void_p(stream.cuda_stream))
    assert rc == 0, lib.WebRtcNsB200_LastError()
    batch = pkg.NsBatch(n, a.fs, a.mode, fixed=a.fixed, devices=[dev.index])

    def tick(i):
        off = i * fl * 2
        batch.process_device(pcm_in.data_ptr() + off, total, pcm_out.data_ptr() + off, total, 1, stream.cuda_stream)

    for i in range(warm):
        tick(i)
    torch.cuda.synchronize()
    # one event pair per launch: the kernel's own duration, without the host's gap between two calls
    # (at ~130 us per tick the Python caller does not always keep the GPU fed; `tick_ms_wall` shows it)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        ev[i][0].record(stream)
        tick(warm + i)
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    per = sorted(x.elapsed_time(y) for x, y in ev)
    avg_ms = sum(per) / steps
    wall_ms = ev[0][0].elapsed_time(ev[-1][1]) / steps
    batch.close()
    peak, which = measured_peak_gbs()
    bsf = algorithmic_bytes(kind, a.fs, 1)
    achieved = bsf * n / (avg_ms * 1e-3) / 1e9
    kname = "nsx_process_kernel" if a.fixed else "nsf_process_kernel"
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": measured_traffic(kname, 1, n) if a.fs == 16000 else None, "peak_source": which, "kernel": kname,
            "workload": "%d streams x 1 frame per launch (10 ms tick), %d launches after %d warm-up ticks; "
                        "state %.0f MB > L2" % (n, steps, warm, n * S_HOT[(kind, a.fs)] / 1e6),
            "bytes_per_stream_frame": bsf, "frames_per_launch": 1, "launch_ms_avg": avg_ms,
            "launch_ms_median": per[len(per) // 2], "launch_ms_p10_p90": [per[steps // 10], per[steps * 9 // 10]],
            "launch_ms_max": per[-1], "tick_ms_wall": wall_ms,
            "value": n * 0.01 / (avg_ms * 1e-3), "value_unit": "audio-s/s"}


def config5(a, pkg, lib, torch, dist, rank, local_rank, world, barrier):
    """BASELINE configs[4] / SURVEY.md 8d(5): --total-streams streams x --seconds at 16 kHz, each rank owning a
    contiguous slice; PCM generated on the device per chunk of F frames, output reduced on the device to a
    per-stream (sum, energy) checksum.  A step is one chunk (generate + process + reduce) of every stream of
    the rank; the timed region is the whole job unless --steps cuts it short (explicit --steps only)."""
    import numpy as np
    from audiosignalprocess_b200.shard import shard_range
    F = a.frames_per_step
    lo, hi = shard_range(a.total_streams, world, rank)
    n = hi - lo
    frames = a.seconds * 100
    if "--steps" in sys.argv:
        frames = min(frames, a.steps * F)
    steps = -(-frames // F)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)
    batch = pkg.NsBatch(n, a.fs, a.mode, fixed=a.fixed, devices=[local_rank])
    pkg.run_generated_job(batch, max(1, a.warmup) * F, F, first_stream=lo, stream=stream)
    batch.reset(a.mode)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.WebRtcNsB200_KernelLaunches()
    pairs = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sums, _ = pkg.run_generated_job(batch, frames, F, first_stream=lo, stream=stream,
                                    on_launch=lambda i, x, y: pairs.append((x, y)))
    e1.record(stream)
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches = lib.WebRtcNsB200_KernelLaunches() - launches0
    ms_total = e0.elapsed_time(e1)
    ms_proc = sum(x.elapsed_time(y) for x, y in pairs)
    t = torch.tensor([ms_total, ms_proc], dtype=torch.float64, device=dev)
    # a checksum of the per-stream checksums, summed over the ranks: the job's fingerprint
    fp = torch.tensor([int(np.bitwise_xor.reduce(sums[:, 0] * 1000003 + sums[:, 1])) & 0x7fffffffffff, int(sums[:, 1].sum() >> 20)],
                      dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fp, op=dist.ReduceOp.SUM)
    if rank == 0:
        kind = "fixed" if a.fixed else "float"
        peak, which = measured_peak_gbs()
        audio_s = a.total_streams * frames * 0.01
        ms_max, ms_proc_max = float(t[0].item()), float(t[1].item())
        bytes_per_sf = algorithmic_bytes(kind, a.fs, F)
        avg_launch_ms = ms_proc / len(pairs)
        achieved = bytes_per_sf * n * F / (avg_launch_ms * 1e-3) / 1e9
        line = {"metric": "ns_audio_seconds_per_second", "value": audio_s / (ms_max * 1e-3), "unit": "audio-s/s",
                "n_gpus": world, "steps": steps, "warmup": max(1, a.warmup), "ms_per_step": ms_max / steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "int16/int32" if a.fixed else "f32", "data": "synthetic",
                "config": {"workload": "config 5: %d streams x %d s at %d Hz sharded over %d GPU(s) (%d per GPU), %s policy %d, "
                                       "PCM generated on the device per chunk of F=%d frames, output reduced to per-stream checksums"
                                       % (a.total_streams, frames // 100, a.fs, world, n, "WebRtcNsx fixed" if a.fixed else "WebRtcNs float",
                                          a.mode, F),
                           "streams_per_gpu": n, "fs": a.fs, "policy": a.mode, "frames_per_launch": F, "pcm": "int16",
                           "l2": "every step generates and processes fresh PCM (in+out %.0f MB per step) larger than the 126 MB L2"
                                 % (2 * n * F * (a.fs // 100) * 2 / 1e6)},
                "clocks": sampler.summary(), "e2e": None, "gpu_launches": int(launches),
                "process_only_value": audio_s / (ms_proc_max * 1e-3),
                "job_fingerprint": [int(fp[0].item()), int(fp[1].item())],
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "peak_source": which,
                             "kernel": "nsx_process_kernel" if a.fixed else "nsf_process_kernel",
                             "bytes_per_stream_frame": bytes_per_sf, "frames_per_launch": F, "launch_ms_avg": avg_launch_ms},
                "cpu_baseline": None}
        print(json.dumps(line))
    batch.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="streams per GPU")
    ap.add_argument("--fs", type=int, default=16000)
    ap.add_argument("--mode", type=int, default=2)
    ap.add_argument("--fixed", action="store_true", help="WebRtcNsx (fixed point) instead of float NS")
    ap.add_argument("--frames-per-step", type=int, default=100)
    ap.add_argument("--config5", action="store_true",
                    help="BASELINE configs[4]: --total-streams x --seconds of 16 kHz PCM generated on the device chunk "
                         "by chunk, sharded over the ranks, output reduced to per-stream checksums (strong scaling)")
    ap.add_argument("--total-streams", type=int, default=65536)
    ap.add_argument("--seconds", type=int, default=600)
    ap.add_argument("--tick-streams", type=int, default=32768,
                    help="streams of the one-frame-per-launch leg (roofline_tick; rank 0 at N=1 only)")
    ap.add_argument("--no-tick", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the NSx 16/8 kHz x 8192 and 48 kHz x 2048 legs of the N=1 line")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kind = "fixed" if a.fixed else "float"
    F = a.frames_per_step
    fl = a.fs // 100
    workload = "%d independent %d Hz mono streams/GPU, %s policy %d, F=%d frames per launch (%.1f s of audio per step)" % (
        a.streams, a.fs, "WebRtcNsx fixed" if a.fixed else "WebRtcNs float", a.mode, F, F * 0.01)
    config = {"workload": workload, "streams_per_gpu": a.streams, "fs": a.fs, "policy": a.mode,
              "frames_per_launch": F, "pcm": "int16",
              "l2": "every step streams fresh PCM (in+out %.0f MB per step) larger than the 126 MB L2" % (
                  2 * a.streams * F * fl * 2 / 1e6)}

    if a.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        # the GPU arm's own workload: every stream, persistent handles, the same warm-up, the same frames
        value, done = cpu_reference_steps(kind, a.fs, a.mode, a.streams, F, a.steps, a.warmup, cores)
        if value is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libns_ref.so missing"}))
            return 0
        sample = ("%d streams x %d frames per step, %d timed steps after %d warm-up steps on persistent handles, "
                  "unmodified reference C (gcc -O2, oracle/Makefile), %d pthreads taking one stream at a time" % (
                      a.streams, F, done, a.warmup, cores))
        print(json.dumps({
            "impl": "reference", "metric": "ns_audio_seconds_per_second", "value": value, "unit": "audio-s/s",
            "n_gpus": a.gpus, "steps": done, "warmup": a.warmup,
            "ms_per_step": 1e3 * a.streams * F * 0.01 / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if not a.fixed else "int16/int32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import torch.distributed as dist
    import audiosignalprocess_b200 as pkg

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = pkg.load_library()
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if a.config5:
        return config5(a, pkg, lib, torch, dist, rank, local_rank, world, barrier)

    total_steps = a.warmup + a.steps
    n_samples = total_steps * F * fl
    pcm_in = torch.empty((a.streams, n_samples), dtype=torch.int16, device=dev)
    pcm_out = torch.empty_like(pcm_in)
    # a real (non-default) stream: the library treats NULL as "use my own stream", and the
    # timing events must sit on the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    rc = lib.WebRtcNsB200_SynthPcmDevice(C.c_void_p(pcm_in.data_ptr()), n_samples, a.streams, rank * a.streams,
                                         a.fs, 0, n_samples, 1234, C.c_void_p(stream.cuda_stream))
    assert rc == 0, lib.WebRtcNsB200_LastError()
    torch.cuda.synchronize()
    batch = pkg.NsBatch(a.streams, a.fs, a.mode, fixed=a.fixed, devices=[local_rank])

    def step_device(i):
        off = i * F * fl * 2
        batch.process_device(pcm_in.data_ptr() + off, n_samples, pcm_out.data_ptr() + off, n_samples, F,
                             stream.cuda_stream)

    for i in range(a.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.WebRtcNsB200_KernelLaunches()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record(stream)
    for i in range(a.steps):
        step_device(a.warmup + i)
        ev[i + 1].record(stream)
    barrier()
    launches = lib.WebRtcNsB200_KernelLaunches() - launches0
    ms_total = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    audio_s = a.streams * a.steps * F * 0.01 * world
    value = audio_s / (ms_total_max * 1e-3)

    # ---- end to end through the host-pointer C calls, pinned buffers.  Every step copies its input from
    # pinned host memory and its result is read on the host.  Two forms of the same entry point:
    #   blocking   one WebRtcNs_ProcessBatch per step (each call pays its own pipeline fill and drain)
    #   streaming  WebRtcNs_ProcessBatchAsync with two steps in flight on two buffer pairs, a step's result
    #              read after WebRtcNsB200_WaitBatch -- how a caller that streams audio uses the library
    e2e = None
    if not a.no_e2e:
        e_steps = a.steps
        n1 = F * fl
        host_in = [torch.empty((a.streams, n1), dtype=torch.int16).pin_memory() for _ in range(2)]
        host_out = [torch.empty((a.streams, n1), dtype=torch.int16).pin_memory() for _ in range(2)]
        for k in range(2):
            host_in[k].copy_(pcm_in[:, k * n1:(k + 1) * n1].cpu())

        def timed(fn):
            batch.reset(a.mode)
            fn(a.warmup)
            barrier()
            t0 = time.perf_counter()
            fn(e_steps)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return a.streams * e_steps * F * 0.01 * world / float(tt.item())

        def run_blocking(steps):
            for i in range(steps):
                k = i & 1
                batch.process_ptr(host_in[k].data_ptr(), n1, host_out[k].data_ptr(), n1, F)
                _ = int(host_out[k][0, 0])   # the step's result is read on the host

        def run_streaming(steps):
            tickets = [None, None]
            for i in range(steps):
                k = i & 1
                if tickets[k] is not None:
                    batch.wait(tickets[k])
                    _ = int(host_out[k][0, 0])   # result of step i-2, read before its buffers are reused
                tickets[k] = batch.process_ptr_async(host_in[k].data_ptr(), n1, host_out[k].data_ptr(), n1, F)
            for k in ((steps & 1), 1 - (steps & 1)):
                if tickets[k] is not None:
                    batch.wait(tickets[k])
                    _ = int(host_out[k][0, 0])

        v_block = timed(run_blocking)
        v_stream = timed(run_streaming)
        # what the copies could do at best: every rank probes its own GPU's H2D + D2H duplex rate at the same
        # time (ranks share the host's memory / PCIe fabric); the end-to-end path moves bytes_per_step each way
        barrier()
        probe = pcie_probe(torch, dev)
        pt = torch.tensor([probe], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(pt, op=dist.ReduceOp.SUM)
        probe_total = float(pt.item())
        bytes_dir = a.streams * n1 * 2
        gbs_dir = v_stream / (a.streams * F * 0.01 * world) * bytes_dir * world / 1e9   # steps/s x bytes per step, all ranks
        e2e = {"value": v_stream, "unit": "audio-s/s",
               "h2d_bytes_per_step": bytes_dir, "d2h_bytes_per_step": bytes_dir,
               "api": "WebRtcNs%s_ProcessBatchAsync + WebRtcNsB200_WaitBatch, two steps in flight" % ("x" if a.fixed else ""),
               "blocking_value": v_block, "blocking_api": "WebRtcNs%s_ProcessBatch, one call per step" % ("x" if a.fixed else ""),
               "pcie_gbs": gbs_dir, "pcie_probe_gbs": probe_total,
               "pcie_note": "GB/s per direction summed over the %d rank(s), both directions busy at once: what the end-to-end "
                            "leg moved, and what bare pinned-buffer H2D + D2H copy pairs sustain for one second on all ranks "
                            "at the same time (the ranks share the host's fabric: profiles/r2_pcie_probe_multi.log)" % world,
               "frac_of_pcie_probe": gbs_dir / probe_total if probe_total > 0 else None}
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank == 0:
        peak, which = measured_peak_gbs()
        bytes_per_sf = algorithmic_bytes(kind, a.fs, F)
        med = sorted(step_ms)[len(step_ms) // 2]
        avg_ms = ms_total / a.steps      # rank 0's own launches
        achieved = bytes_per_sf * a.streams * F / (avg_ms * 1e-3) / 1e9
        kname = "nsx_process_kernel" if a.fixed else "nsf_process_kernel"
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": measured_traffic(kname, F, a.streams) if a.fs == 16000 else None,
                    "traffic_source": _profile_summary(kname, F, a.streams)[1] if a.fs == 16000 else None,
                    "peak_source": which, "kernel": kname,
                    "bytes_per_stream_frame": bytes_per_sf, "frames_per_launch": F,
                    "launch_ms_avg": avg_ms, "launch_ms_median": med}
        # The kernel is issue-bound at large F: the same launch time against one warp
        # instruction per scheduler-cycle (4 schedulers x SMs x the SM clock sampled during the run).
        wi = measured_instructions(kname, F, a.streams) if a.fs == 16000 else None
        clk = sampler.summary().get("sm_mhz")
        if wi and clk:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            roofline["issue"] = {"warp_instructions_per_stream_frame": wi, "source": _profile_summary(kname, F, a.streams)[1],
                                 "frac_of_issue_peak": wi * a.streams * F / (avg_ms * 1e-3) / (4.0 * sms * clk * 1e6)}
        cpu = None
        if not a.no_cpu:
            cores = os.cpu_count() or 1
            # the same workload through the unmodified reference: every stream, persistent handles, the same
            # warm-up steps, then as many of the timed steps as fit ~15 s (the reference runs ~1000 audio-s/s per core)
            c_steps = max(4, min(a.steps, int(15.0 * 1000.0 * cores / (a.streams * F * 0.01 * (3.0 if a.fs > 16000 else 1.0)))))
            v, done = cpu_reference_steps(kind, a.fs, a.mode, a.streams, F, c_steps, a.warmup, cores, first_stream=rank * a.streams)
            if v is not None:
                cpu = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "reference",
                       "sample": "%d streams x %d frames per step, %d timed steps after %d warm-up steps on persistent handles, "
                                 "%d pthreads taking one stream at a time" % (a.streams, F, done, a.warmup, cores)}
        tick = None
        extra = None
        if world == 1 and not a.no_tick and F != 1:
            del pcm_in, pcm_out
            torch.cuda.empty_cache()
            tick = tick_roofline(a, pkg, lib, torch, dev, stream, kind)
        if world == 1 and not a.no_extra and not a.fixed and a.fs == 16000 and F == 100:
            # BASELINE.json configs[2] and configs[3] beside the headline (configs[1]), so that the driver's record holds them
            extra = {"nsx_16k_8192": extra_leg(pkg, lib, torch, dev, stream, "fixed", 16000, 8192, 100, a.mode),
                     "nsx_8k_8192": extra_leg(pkg, lib, torch, dev, stream, "fixed", 8000, 8192, 100, a.mode),
                     "float_48k_2048": extra_leg(pkg, lib, torch, dev, stream, "float", 48000, 2048, 50, a.mode)}
        line = {"metric": "ns_audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_total_max / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int16/int32" if a.fixed else "f32",
                "data": "synthetic", "config": config, "clocks": sampler.summary(), "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "roofline_tick": tick, "extra": extra, "cpu_baseline": cpu}
        print(json.dumps(line))
    batch.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
