"""Per-chunk stage timeline of one 48 kHz device batch (NSB200_TRACE=2): python tools/band_trace.py [chunk_frames] [streams] [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    os.environ["NSB200_BAND_CHUNK"] = sys.argv[1]
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
F = int(sys.argv[3]) if len(sys.argv) > 3 else 50
import torch
import audiosignalprocess_b200 as pkg
fs, fl = 48000, 480
x = torch.from_numpy(pkg.synth_pcm_host(S, fs, F * fl)).cuda()
y = torch.empty_like(x)
b = pkg.NsBatch(S, fs, 2, devices=[0])
st = torch.cuda.Stream()
for i in range(3):
    b.process_device(x.data_ptr(), F * fl, y.data_ptr(), F * fl, F, st.cuda_stream)
torch.cuda.synchronize()
os.environ["NSB200_TRACE"] = "2"
for i in range(2):
    b.process_device(x.data_ptr(), F * fl, y.data_ptr(), F * fl, F, st.cuda_stream)
    torch.cuda.synchronize()
    print("--", file=sys.stderr)
