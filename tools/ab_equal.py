"""Tuning helper (GPU box): do two builds of the library produce the same bits?
  NSB200_LIB=<variant.so> python tools/ab_equal.py dump /tmp/a.npz ; python tools/ab_equal.py dump /tmp/b.npz
  python tools/ab_equal.py cmp /tmp/a.npz /tmp/b.npz
Float NS at 16 kHz (int16 and float I/O, 128 streams x 700 frames in launches of 7 frames), 8 kHz, 48 kHz."""
import sys
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if sys.argv[1] == "dump":
    import audiosignalprocess_b200 as pkg
    out = {}
    for fs, n, frames, fpl in [(16000, 128, 700, 7), (8000, 64, 300, 300), (48000, 16, 100, 50)]:
        fl = fs // 100
        x = pkg.synth_pcm_host(n, fs, frames * fl)
        b = pkg.NsBatch(n, fs, 2)
        o = np.zeros_like(x)
        for f0 in range(0, frames, fpl):
            o[:, f0 * fl:(f0 + fpl) * fl] = b.process(np.ascontiguousarray(x[:, f0 * fl:(f0 + fpl) * fl]))
        b.close()
        out["i16_%d" % fs] = o
    x = pkg.synth_pcm_host(32, 16000, 300 * 160)
    b = pkg.NsBatch(32, 16000, 3)
    out["f32_16000"] = b.process_bands_f32(x.astype(np.float32).reshape(32, 300, 1, 160))
    b.close()
    np.savez(sys.argv[2], **out)
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    ok = True
    for k in a.files:
        same = np.array_equal(a[k], b[k])
        ok &= same
        print(k, "identical" if same else "DIFFERENT (max |d| %g)" % np.abs(a[k].astype(np.float64) - b[k]).max(), a[k].shape)
    sys.exit(0 if ok else 1)
