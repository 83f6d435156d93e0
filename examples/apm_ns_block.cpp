// The C++ mirror of the author's wrapper class (include/apm_ns_b200.h) used exactly as
// WebRtc_AMP_Port/libapm's callers use APM_NS: initNsModule once, processCaptureStream per block,
// interleaved samples processed in place.  Reads raw interleaved PCM from stdin-like files so that the
// test suite can compare it with the reference class (tests/test_gpu_reference_caller.py).
//
//   apm_ns_block <fs> <mode> <channels> <frames_per_call> <i16|f32> <in.raw> <out.raw>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/apm_ns_b200.h"

int main(int argc, char** argv) {
  if (argc != 8) {
    fprintf(stderr, "usage: %s fs mode channels frames_per_call i16|f32 in.raw out.raw\n", argv[0]);
    return 2;
  }
  const unsigned fs = (unsigned)atoi(argv[1]);
  const int mode = atoi(argv[2]), channels = atoi(argv[3]), fpc = atoi(argv[4]);
  const bool f32 = strcmp(argv[5], "f32") == 0;
  const int fl = (int)fs / 100;
  FILE* fi = fopen(argv[6], "rb");
  FILE* fo = fopen(argv[7], "wb");
  if (!fi || !fo) return 3;
  APM_NS ns;
  if (!ns.initNsModule(fs, mode, fl, channels)) {
    fprintf(stderr, "initNsModule failed: %s\n", WebRtcNsB200_LastError());
    return 4;
  }
  const size_t esz = f32 ? sizeof(float) : sizeof(short);
  std::vector<char> buf((size_t)fpc * fl * channels * esz);
  for (;;) {
    const size_t got = fread(buf.data(), esz * channels * fl, (size_t)fpc, fi);   // whole frames
    if (got == 0) break;
    if (f32) ns.processCaptureStream(reinterpret_cast<float*>(buf.data()), (int)got * fl, channels);
    else ns.processCaptureStream(reinterpret_cast<short*>(buf.data()), (int)got * fl, channels);
    fwrite(buf.data(), esz * channels * fl, got, fo);
  }
  fclose(fi);
  fclose(fo);
  return 0;
}
