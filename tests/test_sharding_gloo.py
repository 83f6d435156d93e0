"""N > 1 host logic on CPU: two gloo ranks shard the streams, process their slice and gather on
rank 0.  The per-slice processing is stood in by the CPU oracle (a test-only checker) -- what is
under test is the partitioning and the gather, which are the only multi-GPU code on this path."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

WORKER = r'''
import os, sys, ctypes as C
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import audiosignalprocess_b200 as pkg
from audiosignalprocess_b200.shard import process_sharded, shard_range
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
ora = C.CDLL(os.path.join(sys.argv[1], "oracle", "liboracle_ns.so"))
fs, mode, frames, n = 16000, 2, 60, 5
x = pkg.synth_pcm_host(n, fs, frames * 160)
def nsx(p):
    out = np.zeros_like(p)
    for s in range(p.shape[0]):
        xs = np.ascontiguousarray(p[s]); o = np.zeros_like(xs)
        ora.nsx_oracle_run(fs, mode, frames, xs.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
        out[s] = o
    return out
full = process_sharded(nsx, x, dist)
if dist.get_rank() == 0:
    assert full.shape == x.shape and np.array_equal(full, nsx(x))
    lo0, hi0 = shard_range(n, 2, 0); lo1, hi1 = shard_range(n, 2, 1)
    assert (lo0, hi0, lo1, hi1) == (0, 3, 3, 5)
    print("SHARD_OK")
else:
    assert full is None
dist.barrier()
dist.destroy_process_group()
'''


def test_shard_range_covers_all_streams():
    from audiosignalprocess_b200.shard import shard_range
    for n in (1, 5, 4096, 65536, 7):
        for g in (1, 2, 4, 8):
            spans = [shard_range(n, g, r) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert max(hi - lo for lo, hi in spans) <= -(-n // g)


def test_two_rank_gloo_gather(tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD_OK" in outs[0]
