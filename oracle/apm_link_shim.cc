// TEST INFRASTRUCTURE ONLY.
//
// Boundary proof: the reference's own caller of the path -- the author's class APM_NS
// (WebRtc_AMP_Port/libapm/src/apm_ns.cpp, compiled UNMODIFIED where it lies, together with the reference's
// AudioBuffer / splitting filter / resampler) -- linked against libwebrtc_ns_b200.so INSTEAD of the
// reference's ns/*.o.  apm_ns.cpp calls WebRtcNs_Create / Init / set_policy / Analyze / Process / Free per
// channel and per 10 ms frame (:17-24, :69-74, :113-118); those symbols now resolve to the product library.
// This file only exposes the class through a C interface for the test (tests/test_gpu_reference_caller.py).
#include <stdint.h>

#include "libapm/include/apm_ns.h"

extern "C" {

void* apm_b200_create(unsigned fs, int mode, int frame, int channels) {
  APM_NS* a = new APM_NS();
  if (!a->initNsModule(fs, mode, frame, channels)) {
    delete a;
    return 0;
  }
  return a;
}
void apm_b200_free(void* a) { delete (APM_NS*)a; }
void apm_b200_process_i16(void* a, int16_t* data, int frame, int channels) {
  ((APM_NS*)a)->processCaptureStream(data, frame, channels);
}
void apm_b200_process_f32(void* a, float* data, int frame, int channels) {
  ((APM_NS*)a)->processCaptureStream(data, frame, channels);
}

}  // extern "C"
