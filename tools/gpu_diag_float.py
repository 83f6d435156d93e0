"""GPU diagnostic: where does the float kernel diverge from the reference, and is the
output independent of how frames are chunked into launches?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import audiosignalprocess_b200 as pkg
from conftest import RefLib, snr_db

fs = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
mode = 2
n = 8
fl = fs // 100
ref = RefLib(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libns_ref.so"))
x = pkg.synth_pcm_host(n, fs, frames * fl)
xin = x.astype(np.float32).reshape(n, frames, 1, fl)


def run(chunks):
    b = pkg.NsBatch(n, fs, mode)
    out = np.zeros((n, frames, 1, fl), np.float32)
    f0 = 0
    ci = 0
    while f0 < frames:
        nf = min(chunks[min(ci, len(chunks) - 1)], frames - f0)
        out[:, f0:f0 + nf] = b.process_bands_f32(xin[:, f0:f0 + nf])
        f0 += nf
        ci += 1
    b.close()
    return out.reshape(n, frames * fl)


o1 = run([1])
o2 = run([1, 2, 7, 40, 250, 10 ** 9])
o3 = run([10 ** 9])
print("F=1 vs ragged identical:", np.array_equal(o1, o2), " F=1 vs one launch identical:", np.array_equal(o1, o3))
for s in range(n):
    reff, _, _ = ref.ns(fs, mode, x[s])
    for name, o in (("F=1", o1), ("ragged", o2), ("all", o3)):
        d = np.abs(o[s] - reff).reshape(frames, fl).max(1)
        bad = np.nonzero(d > 0.5)[0]
        print("stream %d %-6s max %.3f snr %.1f dB first frame >0.5: %s  frames>0.5: %d" % (
            s, name, d.max(), snr_db(reff, o[s]), bad[0] if len(bad) else None, len(bad)))
    if not np.array_equal(o1[s], o2[s]):
        dd = np.abs(o1[s] - o2[s]).reshape(frames, fl).max(1)
        print("   F=1 vs ragged first differing frame:", np.nonzero(dd > 0)[0][:5], "max", dd.max())
