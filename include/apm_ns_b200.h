/*
 * C++ mirror of the author's wrapper class APM_NS (WebRtc_AMP_Port/libapm/include/apm_ns.h:14-53,
 * libapm/src/apm_ns.cpp) on top of libwebrtc_ns_b200.so: same method names, arguments and in-place
 * semantics; the per-channel loop, AudioBuffer split/merge and the (de)interleave run on the GPU in
 * one call (WebRtcNs_ProcessInterleaved*).  Header only.
 */
#ifndef APM_NS_B200_H_
#define APM_NS_B200_H_

#include <vector>

#include "webrtc_ns_b200.h"

enum { NS_Mode_Mild = 0, Ns_Mode_Mideum, Ns_Mode_Aggressive };   /* apm_ns.h:8-12 */

class APM_NS {
 public:
  APM_NS() : m_frequency(0), m_channels(0), m_ns_mode(0), init_flag(false) {}
  ~APM_NS() {
    for (size_t i = 0; i < m_handles.size(); ++i) WebRtcNs_Free(m_handles[i]);
  }
  /* apm_ns.cpp:7-45: one handle per channel, Init + set_policy; false on any failure */
  bool initNsModule(unsigned int frequency, int ns_mode, int input_frames, int input_channels) {
    (void)input_frames;
    m_frequency = frequency;
    m_channels = input_channels;
    m_ns_mode = ns_mode;
    if (m_channels < 0) return false;
    for (size_t i = 0; i < m_handles.size(); ++i) WebRtcNs_Free(m_handles[i]);
    m_handles.clear();
    for (int i = 0; i < m_channels; ++i) {
      NsHandle* h = NULL;
      if (WebRtcNs_Create(&h)) return false;
      m_handles.push_back(h);
      if (WebRtcNs_Init(h, frequency)) return false;
      if (WebRtcNs_set_policy(h, ns_mode)) return false;
    }
    init_flag = true;
    return true;
  }
  /* apm_ns.cpp:47-89: interleaved float in [-1, 1], processed in place */
  void processCaptureStream(float* data, int samples_per_channel, int input_channels) {
    if (!init_flag || input_channels != m_channels) return;
    WebRtcNs_ProcessInterleavedF32(m_handles.data(), m_channels, data, samples_per_channel);
  }
  /* apm_ns.cpp:91-132: interleaved int16, processed in place */
  void processCaptureStream(short* data, int samples_per_channel, int input_channels) {
    if (!init_flag || input_channels != m_channels) return;
    WebRtcNs_ProcessInterleavedI16(m_handles.data(), m_channels, data, samples_per_channel);
  }

 private:
  APM_NS(const APM_NS&);
  APM_NS& operator=(const APM_NS&);
  std::vector<NsHandle*> m_handles;
  unsigned int m_frequency;
  int m_channels;
  int m_ns_mode;
  bool init_flag;
};

#endif /* APM_NS_B200_H_ */
