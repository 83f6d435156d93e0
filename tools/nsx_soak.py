#!/usr/bin/env python
"""NSx bit-exactness soak (GPU box): 256 streams x 60 s at every rate and policy, the GPU library against the
compiled reference (oracle/_ref/libns_ref.so), in launches of ragged length.  Test infrastructure.

  python tools/nsx_soak.py [--streams 256] [--seconds 60] [--out profiles/r2_nsx_soak.md]"""
import argparse
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import RefLib  # noqa: E402

import audiosignalprocess_b200 as pkg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=256)
    ap.add_argument("--seconds", type=int, default=60)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    ref = RefLib(os.path.join(ROOT, "oracle", "_ref", "libns_ref.so"))
    rows, bad = [], 0
    t0 = time.time()
    for fs in (8000, 16000, 32000, 48000):
        for mode in (0, 1, 2, 3):
            n, frames, fl = a.streams, a.seconds * 100, fs // 100
            x = pkg.synth_pcm_host(n, fs, frames * fl, base_seed=4242 + mode)
            b = pkg.NsBatch(n, fs, mode, fixed=True)
            out = np.zeros_like(x)
            f0, ci = 0, 0
            chunks = [1, 3, 50, 333, 1000]
            while f0 < frames:
                nf = min(chunks[ci % len(chunks)], frames - f0)
                out[:, f0 * fl:(f0 + nf) * fl] = b.process(np.ascontiguousarray(x[:, f0 * fl:(f0 + nf) * fl]))
                f0 += nf
                ci += 1
            b.close()
            with ThreadPoolExecutor(os.cpu_count() or 8) as ex:
                refs = list(ex.map(lambda s: ref.nsx(fs, mode, x[s]), range(n)))
            diff = sum(int(not np.array_equal(refs[s], out[s])) for s in range(n))
            bad += diff
            rows.append("| %d | %d | %d x %d s | %d |" % (fs, mode, n, a.seconds, diff))
            print(rows[-1], flush=True)
    text = ("NSx soak: GPU library vs the compiled reference, int16 output compared bit for bit (`tools/nsx_soak.py`, "
            "%.0f s)\n\n| fs | policy | streams x length | streams that differ |\n|---|---|---|---|\n" % (time.time() - t0)) + "\n".join(rows) + "\n"
    if a.out:
        open(a.out, "w").write(text)
    print("TOTAL differing streams:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
