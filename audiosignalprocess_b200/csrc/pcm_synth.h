// Integer-only, random-access synthetic PCM (SURVEY.md section 8d "Synthetic
// inputs"): sample n of stream s is a pure function of (base_seed, s, fs, n), so
// the C oracle, the CPU baseline and the on-device generator produce identical
// int16 without sharing state.  Compiles as C99, C++ and CUDA.
//
// Stream classes, cycled by stream_idx % 8:
//   0 white noise (-30 dBFS)            1 pink-ish noise (octave-held sum, -30 dBFS)
//   2 white + 440 Hz / 1 kHz tone bursts (0.5 s on / 1.5 s off, -12 dBFS)
//   3 pink + linear chirp bursts 300->3400 Hz
//   4 white with 2 s of digital silence starting at t = 1.5 s   (zero-energy path)
//   5 white + tone with full-scale clipping bursts               (saturation path)
//   6 digital silence for the first 1 s, then white + tone      (delayed start-up)
//   7 speech-like: noise amplitude-modulated at a syllabic 4 Hz
#ifndef AUDIOSIGNALPROCESS_B200_PCM_SYNTH_H_
#define AUDIOSIGNALPROCESS_B200_PCM_SYNTH_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define PCM_SYNTH_FN __host__ __device__ static __forceinline__
#else
#define PCM_SYNTH_FN static inline
#endif

PCM_SYNTH_FN uint32_t pcm_mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

// Uniform in [-32768, 32767], counter based.
PCM_SYNTH_FN int32_t pcm_white16(uint32_t seed, uint32_t n) {
  uint32_t h = pcm_mix32(seed * 0x9E3779B9u + pcm_mix32(n + 0x85EBCA6Bu));
  return (int32_t)(h >> 16) - 32768;
}

// Parabolic sine, phase in 1/65536 turns, output about +-32767 (a few % THD:
// it is a test stimulus, not a reference oscillator).
PCM_SYNTH_FN int32_t pcm_sin16(uint32_t phase) {
  int32_t x = (int32_t)(phase & 0xFFFFu);     // 0..65535 = one turn
  int32_t half = x & 0x7FFF;                  // position in the half turn, 0..32767
  int32_t y = (half * (32768 - half)) >> 13;  // 0..32768 parabola
  if (y > 32767) y = 32767;
  return (x & 0x8000) ? -y : y;
}

PCM_SYNTH_FN int16_t pcm_sat16(int32_t v) {
  return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
}

// One sample. fs in Hz, n = sample index from stream start.
PCM_SYNTH_FN int16_t pcm_synth_sample(uint32_t base_seed, uint32_t stream_idx, uint32_t fs,
                                      uint32_t n) {
  const uint32_t seed = base_seed + stream_idx;
  const uint32_t cls = stream_idx & 7u;
  const uint32_t ms = (uint32_t)(((uint64_t)n * 1000u) / fs);  // time in ms
  // -30 dBFS RMS for a uniform variable: peak = 32768 * 10^(-30/20) * sqrt(3) ~ 1795
  int32_t white = (pcm_white16(seed, n) * 1795) >> 15;
  int32_t pink = 0;
  {
    int k;
    for (k = 0; k < 6; ++k) pink += pcm_white16(seed ^ (0x51ED27u * (uint32_t)(k + 1)), n >> k);
    pink = (pink * 733) >> 15;  // six held octaves, scaled back to about -30 dBFS
  }
  const uint32_t burst_ms = ms % 2000u;
  const int in_burst = burst_ms < 500u;
  const uint32_t burst_no = ms / 2000u;
  int32_t v = 0;
  switch (cls) {
    case 0: v = white; break;
    case 1: v = pink; break;
    case 2:
    case 5:
    case 6: {
      v = white;
      if (in_burst) {
        // 440 Hz and 1 kHz alternate per burst; phase step = f * 65536 / fs
        uint32_t f = (burst_no & 1u) ? 1000u : 440u;
        uint32_t ph = (uint32_t)(((uint64_t)n * f * 65536u) / fs);
        v += (pcm_sin16(ph) * 8231) >> 15;  // -12 dBFS peak
      }
      if (cls == 5 && burst_ms >= 1000u && burst_ms < 1200u) v *= 40;  // clip
      if (cls == 6 && ms < 1000u) v = 0;
      break;
    }
    case 3: {
      v = pink;
      if (in_burst) {
        // linear chirp 300 -> 3400 Hz across the 0.5 s burst; phase is the
        // integral of the instantaneous frequency, in integer arithmetic.
        uint64_t t = (uint64_t)n - (uint64_t)burst_no * 2u * fs;  // samples into the burst
        uint64_t len = fs / 2u;
        uint64_t cyc16 = (300u * t * 65536u) / fs + (((3100u * t * t) / len) * 32768u) / fs;
        v += (pcm_sin16((uint32_t)cyc16) * 8231) >> 15;
      }
      break;
    }
    case 4: v = (ms >= 1500u && ms < 3500u) ? 0 : white; break;
    default: {  // 7: 4 Hz triangular envelope on 4x louder noise
      uint32_t p = ms % 250u;
      int32_t env = (int32_t)(p < 125u ? p : 250u - p);  // 0..125
      v = (white * 4 * env) / 125;
      break;
    }
  }
  return pcm_sat16(v);
}

#endif  // AUDIOSIGNALPROCESS_B200_PCM_SYNTH_H_
