// Band-split state slab (one per stream handle), 32-bit words.  Plain C++.
#ifndef AUDIOSIGNALPROCESS_B200_BAND_LAYOUT_H_
#define AUDIOSIGNALPROCESS_B200_BAND_LAYOUT_H_

#include <stdint.h>

namespace nsb200 {

enum : int {
  // TwoBandsStates x 3 (splitting_filter.h:33-46, splitting_filter.cc:24-26): each is
  // analysis_state1[6] | analysis_state2[6] | synthesis_state1[6] | synthesis_state2[6]
  kBandOffQmf0 = 0,    // two_bands_states_: the 32 kHz split, or the first stage at 48 kHz
  kBandOffQmf1 = 24,   // band1_states_: 0-16 kHz half -> bands 0, 1
  kBandOffQmf2 = 48,   // band2_states_: 16-32 kHz half -> (dropped), band 2
  kBandOffAnaHist = 72,    // 64 int16: last input samples of the 480 -> 640 resampler
  kBandOffSynHist = 104,   // 64 int16: last input samples of the 640 -> 480 resampler
  kBandStateWords = 160,
};

}  // namespace nsb200

#endif  // AUDIOSIGNALPROCESS_B200_BAND_LAYOUT_H_
