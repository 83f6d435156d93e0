// Host-side constants of the 48 kHz band split: the windowed-sinc kernel tables of
// SincResampler::InitializeKernel (common_audio/resampler/sinc_resampler.cc:208-242) for the
// two ratios used by SplittingFilter (splitting_filter.cc:27-30), and the state the
// 640 -> 480 resampler is left in by PushSincResampler's zero-primed first pass
// (push_sinc_resampler.cc:64-73,85-91).  Plain C++, no CUDA calls.
#ifndef AUDIOSIGNALPROCESS_B200_BAND_HOST_INIT_H_
#define AUDIOSIGNALPROCESS_B200_BAND_HOST_INIT_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "band_layout.h"

namespace nsb200 {

// kernel[offset_idx * 32 + i], offset_idx 0..32, i 0..31; same float/double mix as the reference.
inline void band_make_sinc_kernel(double io_ratio, float* kernel) {
  const double kPi = 3.14159265358979323846;
  const double kAlpha = 0.16, kA0 = 0.5 * (1.0 - kAlpha), kA1 = 0.5, kA2 = 0.5 * kAlpha;
  double scale = io_ratio > 1.0 ? 1.0 / io_ratio : 1.0;
  scale *= 0.9;
  for (int o = 0; o <= 32; ++o) {
    const float sub = (float)o / 32;
    for (int i = 0; i < 32; ++i) {
      const float pre_sinc = (float)(kPi * (i - 32 / 2 - sub));
      const float x = (i - sub) / 32;
      const float window = (float)(kA0 - kA1 * cos(2.0 * kPi * x) + kA2 * cos(4.0 * kPi * x));
      kernel[o * 32 + i] = (float)(window * ((pre_sinc == 0) ? scale : (sin(scale * pre_sinc) / pre_sinc)));
    }
  }
}

// virtual_source_idx_ after the first (primed, discarded) pass of a PushSincResampler(640, 480):
// request 640, block 624, ChunkSize() outputs from an all-zero buffer, positions advanced by
// repeated double addition exactly as sinc_resampler.cc:283-326 does.
inline double band_down_initial_vsi() {
  const double r = 640.0 / 480.0;
  const double block = 640 - 32 / 2;           // r4 - r2 with r0 = buffer + 16 (sinc_resampler.cc:196-199)
  double vsi = 0.0;
  int remaining = (int)(block / r);            // ChunkSize(), :346-348
  while (remaining) {
    for (int i = (int)ceil((block - vsi) / r); i > 0; --i) {
      vsi += r;
      if (!--remaining) return vsi;
    }
    vsi -= block;
  }
  return vsi;
}

// Replays SincResampler::Resample's block loop (sinc_resampler.cc:269-342) of a 640 -> 480
// PushSincResampler for `frames` consecutive calls starting from the running position *vsi
// (in our coordinates: block of 640, see band_init_vsi), and records for every output where
// the 32-tap window starts in the kernel's E buffer ([64 history | 640 new]: old block at
// q - 608, new block at q + 32), the table row and the two interpolation weights exactly as
// the reference casts them (sinc_resampler_sse.cc:43,45).
inline void band_down_schedule(double* vsi, int frames, int32_t* sched) {
  const double r = 640.0 / 480.0;
  const double block = 640.0;
  double v = *vsi;
  for (int f = 0; f < frames; ++f) {
    int32_t* out = sched + (size_t)f * 480 * 3;
    int n = 0, remaining = 480;
    bool shifted = false;
    while (remaining) {
      int cnt = (int)ceil((block - v) / r);
      for (; cnt > 0 && remaining; --cnt) {
        const int sidx = (int)v;
        const double voff = (v - sidx) * 32.0;
        const int off = (int)voff;
        const double kif = voff - off;
        const float f2 = (float)kif, f1 = (float)(1.0 - kif);
        out[3 * n] = ((shifted ? sidx + 32 : sidx - 608) << 8) | off;
        memcpy(&out[3 * n + 1], &f2, 4);
        memcpy(&out[3 * n + 2], &f1, 4);
        v += r;
        ++n;
        --remaining;
      }
      if (!remaining) break;
      v -= block;
      shifted = true;
    }
  }
  *vsi = v;
}

// After the primed pass the reference still works on its 624-sample first block; we always
// work on 640-sample blocks whose start lies 16 samples later, so the carried position is
// shifted by the same 16 samples (exactly representable; the samples in between are the zero
// priming).
inline double band_init_vsi() { return band_down_initial_vsi() + 16.0; }

inline void band_init_state(uint32_t* slab) {
  memset(slab, 0, sizeof(uint32_t) * (size_t)kBandStateWords);
}

}  // namespace nsb200

#endif  // AUDIOSIGNALPROCESS_B200_BAND_HOST_INIT_H_
