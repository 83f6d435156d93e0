"""Timeline of one host-pointer batch call (NSB200_TRACE=1): python tools/e2e_trace.py [chunk_frames] [calls] [trace=1]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and int(sys.argv[1]) > 0:
    os.environ["NSB200_CHUNK_FRAMES"] = sys.argv[1]
ncalls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
import torch
import audiosignalprocess_b200 as pkg
S, F, fl = 4096, 100, 160
x = torch.from_numpy(pkg.synth_pcm_host(S, 16000, F * fl)).pin_memory()
y = torch.empty_like(x).pin_memory()
b = pkg.NsBatch(S, 16000, 2, devices=[0])
for i in range(3):
    b.process_ptr(x.data_ptr(), F * fl, y.data_ptr(), F * fl, F)
if len(sys.argv) <= 3 or sys.argv[3] != "0":
    os.environ["NSB200_TRACE"] = "1"
for i in range(ncalls):
    t0 = time.perf_counter()
    b.process_ptr(x.data_ptr(), F * fl, y.data_ptr(), F * fl, F)
    print("call %d: %.3f ms wall" % (i, 1e3 * (time.perf_counter() - t0)), file=sys.stderr)
