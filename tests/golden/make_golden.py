"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libns_ref.so, built from
/root/reference by oracle/Makefile).  Run in the container that has the reference:

    python tests/golden/make_golden.py

Each file holds checksums of the synthetic input (csrc/pcm_synth.h regenerates it) and what the
reference produced for it: int16 output of WebRtcNsx_Process, float output of
WebRtcNs_Analyze+Process and the prior speech probability per frame."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import audiosignalprocess_b200 as pkg  # noqa: E402
from conftest import RefLib  # noqa: E402

SEED = 4242
STREAMS = [3, 4, 5]   # pink + chirp bursts, mid-stream digital silence, clipping bursts


def main():
    ref = RefLib(os.path.join(ROOT, "oracle", "_ref", "libns_ref.so"))
    for fs, mode, frames in ((16000, 2, 220), (8000, 2, 220), (16000, 0, 80), (8000, 3, 80)):
        fl = fs // 100
        x = np.stack([pkg.synth_pcm_host(1, fs, frames * fl, base_seed=SEED, first_stream=s)[0] for s in STREAMS])
        nsx = np.stack([ref.nsx(fs, mode, x[i]) for i in range(len(STREAMS))])
        fl_out, probs = [], []
        for i in range(len(STREAMS)):
            of, _, pp = ref.ns(fs, mode, x[i])
            fl_out.append(of)
            probs.append(pp)
        path = os.path.join(HERE, "ns_fs%d_mode%d.npz" % (fs, mode))
        np.savez_compressed(path, fs=fs, mode=mode, seed=SEED, streams=np.array(STREAMS), pcm_in_sum=x.astype(np.int64).sum(1),
                            pcm_in_sqsum=(x.astype(np.int64) ** 2).sum(1),
                            nsx_out=nsx, ns_out=np.stack(fl_out).astype(np.float32),
                            ns_prior_prob=np.stack(probs).astype(np.float32))
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
