"""bench.py's line contract, as far as it can be checked without a GPU: the reference arm
(`--impl reference`, the unmodified reference C on the host cores) prints ONE JSON line with the
metric, config and the cpu_baseline / e2e objects the driver reads; and the byte accounting of the
roofline is what DESIGN.md states.  The GPU arm of the contract is exercised by the
driver itself."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_match_the_design():
    b = _bench_module()
    # B(F) = IO + 2 * S_hot / F, float 16 kHz: IO 640 B, S_hot 7120 B
    assert b.algorithmic_bytes("float", 16000, 1) == 640 + 2 * 7120
    assert abs(b.algorithmic_bytes("float", 16000, 100) - 782.4) < 1e-9
    assert b.algorithmic_bytes("fixed", 8000, 1) == 320 + 2 * 2302
    assert b.algorithmic_bytes("float", 48000, 1) == 1920 + 2 * 8448


def test_committed_ncu_summaries_feed_the_roofline():
    """roofline.traffic and roofline.issue come from the committed ncu summaries of the same configuration."""
    b = _bench_module()
    t = b.measured_traffic("nsf_process_kernel", 100, 4096)
    assert t is not None and 782 * 409600 * 0.9 < t < 782 * 409600 * 1.5   # close to the algorithmic bytes
    t1 = b.measured_traffic("nsf_process_kernel", 1, 32768)
    assert t1 is not None and 0.8 * 14880 * 32768 < t1 < 1.1 * 14880 * 32768
    wi = b.measured_instructions("nsf_process_kernel", 100, 4096)
    assert wi is not None and 1500 < wi < 4000
    assert b.measured_traffic("nsf_process_kernel", 100, 1234) is None    # another batch size: no claim


@pytest.mark.timeout(300)
def test_reference_arm_prints_the_contract_line():
    ref = os.path.join(ROOT, "oracle", "_ref", "libns_ref.so")
    if not os.path.exists(ref):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"], stdout=subprocess.DEVNULL)
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--streams", "8", "--frames-per-step", "20"],
                         capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ns_audio_seconds_per_second" and d["unit"] == "audio-s/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["streams_per_gpu"] == 8 and d["steps"] == 1 and d["warmup"] == 0
    assert "8 streams x 20 frames per step" in d["cpu_baseline"]["sample"]


@pytest.mark.timeout(300)
def test_reference_arm_never_loads_the_product_library():
    """The CPU arm times oracle/_ref alone: its PCM comes from csrc/pcm_synth.h compiled into the shim, so the
    process that prints the reference line has libns_ref.so mapped and libwebrtc_ns_b200.so not."""
    ref = os.path.join(ROOT, "oracle", "_ref", "libns_ref.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--streams', '4', '--frames-per-step', '10']\n"
            "try:\n    runpy.run_path(%r, run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "maps = open('/proc/self/maps').read()\n"
            "print('MAPS', 'libns_ref.so' in maps, 'libwebrtc_ns_b200' in maps, 'audiosignalprocess_b200' in sys.modules)\n"
            % os.path.join(ROOT, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "MAPS True False False" in out.stdout, out.stdout[-500:]


def test_persistent_reference_batch_matches_one_shot_runs():
    """ref_batch_* (persistent handles + thread pool, the CPU arm) produces what ref_ns_run / ref_nsx_run produce
    stream by stream, across several steps, at 16 and 48 kHz."""
    import ctypes as C
    import numpy as np
    b = _bench_module()
    lib = b.ref_lib()
    if lib is None:
        pytest.skip("oracle/_ref not built")
    for fixed, fs in ((0, 16000), (1, 16000), (0, 48000)):
        fl, n, F, steps = fs // 100, 5, 30, 3
        x = np.zeros((n, steps * F * fl), np.int16)
        lib.ref_synth_pcm(x.ctypes.data, x.shape[1], n, 0, fs, 0, x.shape[1], 99, 2)
        out = np.zeros_like(x)
        h = lib.ref_batch_create(fixed, fs, 2, n, 3)
        assert h
        for k in range(steps):
            xin = np.ascontiguousarray(x[:, k * F * fl:(k + 1) * F * fl])
            o = np.zeros_like(xin)
            assert lib.ref_batch_step(h, F, xin.ctypes.data, F * fl, o.ctypes.data, F * fl) > 0
            out[:, k * F * fl:(k + 1) * F * fl] = o
        lib.ref_batch_free(h)
        for s in range(n):
            want = np.zeros(x.shape[1], np.int16)
            xs = np.ascontiguousarray(x[s])
            if fixed:
                lib.ref_nsx_run(fs, 2, steps * F, xs.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
            else:
                lib.ref_ns_run(fs, 2, steps * F, xs.ctypes.data_as(C.c_void_p), None, want.ctypes.data_as(C.c_void_p), None)
            assert np.array_equal(out[s], want), (fixed, fs, s)
