// Host-side construction of the NSx tables and of a freshly initialised state
// slab (WebRtcNsx_InitCore, ns/nsx_core.c:630-783; WebRtcNsx_set_policy_core,
// :785-813).  Plain C++, no CUDA calls.
//
// The reference carries these tables as literals; every one of them except the
// 17-entry indicator map follows a closed form, which is what we compute here
// (tests/test_tables.py checks each against the reference's literal table):
//   kSinTable1024        trunc(32767 sin(2 pi i/1024))          SPL/complex_fft_tables.h
//   kBlocks*x windows    round(16384 * hybrid Hann)             nsx_core.c:74,90
//   kLogTableFrac        round(256 log2(1 + i/256))             :49
//   kCounterDiv          round(32768/(i+1)), [0] = 32767        :32
//   kLogTable            round(i 256 ln 2)                      :28
//   kLogIndex            round(4096 log2 i)                     :268
//   kFactor1Table        trunc(8192 f1(sqrt(i/256)))            :135 (formula in its comment)
//   kFactor2Aggressiveness1..3  trunc(8192 f2(sqrt(i/256)))     :170,193,216
//   kSumLogIndex, kSumSquareLogIndex, kDeterminantEstMatrix     :240,254,290
#ifndef AUDIOSIGNALPROCESS_B200_NSX_HOST_INIT_H_
#define AUDIOSIGNALPROCESS_B200_NSX_HOST_INIT_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "nsx_layout.h"

namespace nsb200 {

inline int nsx_rnd(double x) { return (int)floor(x + 0.5); }

// FFT twiddles as (cos, sin) int pairs regrouped per stage (layout in ns_fixed.cuh, "Twiddles"): tw[t] is the
// packed 256-point table, n the transform length (256 / 128 complex points).
inline void nsx_fill_fft_twiddles(const uint32_t* tw, int n, int32_t* out) {
  auto put = [&](int idx, int t) {
    out[2 * idx] = (int32_t)(int16_t)(tw[t] & 0xffffu);
    out[2 * idx + 1] = (int32_t)(int16_t)(tw[t] >> 16);
  };
  for (int q = 0; q < 4; ++q) put(q, 32 * q);                                   // stages 0-2
  for (int l = 0; l < 8; ++l) {                                                 // stages 3-5
    put(4 + l, l << 4);
    for (int b = 0; b < 2; ++b) put(4 + 8 * (1 + b) + l, (l + 8 * b) << 3);
    for (int b = 0; b < 4; ++b) put(4 + 8 * (3 + b) + l, (l + 8 * b) << 2);
  }
  if (n == 256) {
    for (int b = 0; b < 2; ++b)
      for (int l = 0; l < 32; ++l) put(60 + b * 32 + l, (l + 32 * b) << 1);     // stage 6
    for (int t = 0; t < 128; ++t) put(124 + t, t);                              // stage 7
  } else {
    for (int b = 0; b < 4; ++b)
      for (int l = 0; l < 16; ++l) put(60 + b * 16 + l, (l + 16 * b) << 1);     // stage 6
  }
}

inline void nsx_fill_tables(NsxTables* t) {
  memset(t, 0, sizeof(*t));
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < 256; ++i) {
    const double v = i < 96 ? sin(pi * i / 192.0) : (i <= 160 ? 1.0 : sin(pi * (256 - i) / 192.0));
    t->win256[i] = (int16_t)nsx_rnd(16384.0 * v);
  }
  for (int i = 0; i < 128; ++i) {
    const double v = i < 48 ? sin(pi * i / 96.0) : (i <= 80 ? 1.0 : sin(pi * (128 - i) / 96.0));
    t->win128[i] = (int16_t)nsx_rnd(16384.0 * v);
  }
  auto sin1024 = [&](int i) -> int {  // first quadrant by formula, the rest by symmetry
    int sign = 1;
    i &= 1023;
    if (i >= 512) { i -= 512; sign = -1; }
    if (i > 256) i = 512 - i;
    return sign * (int)(32767.0 * sin(2.0 * pi * i / 1024.0));
  };
  for (int k = 0; k < 128; ++k) {
    const int c = sin1024(4 * k + 256), s = sin1024(4 * k);
    t->tw[k] = ((uint32_t)c & 0xffffu) | ((uint32_t)s << 16);
  }
  for (int i = 0; i < 256; ++i) t->log_frac[i] = (int16_t)nsx_rnd(256.0 * log2(1.0 + i / 256.0));
  t->counter_div[0] = 32767;
  for (int i = 1; i < 201; ++i) t->counter_div[i] = (int16_t)nsx_rnd(32768.0 / (i + 1));
  for (int i = 0; i < 9; ++i) t->log_tab[i] = (int16_t)nsx_rnd(i * 256.0 * log(2.0));
  for (int i = 1; i < 129; ++i) t->log_idx[i] = (int16_t)nsx_rnd(4096.0 * log2((double)i));
  for (int i = 0; i < 257; ++i) {
    const double g = sqrt(i / 256.0);
    double f = 1.0;
    if (g > 0.5) {
      f = 1.0 + 1.3 * (g - 0.5);
      if (g * f > 1.0) f = 1.0 / g;
    }
    t->factor1[i] = (int16_t)(int)(8192.0 * f);
    for (int k = 0; k < 3; ++k) {
      const double bound = k == 0 ? 0.25 : (k == 1 ? 0.125 : 0.09);
      double gg = g, f2 = 1.0;
      if (g <= 0.5) {
        if (gg <= bound) gg = bound;
        f2 = 1.0 - 0.3 * (0.5 - gg);
      }
      t->factor2[k][i] = (int16_t)(int)(8192.0 * f2);
    }
  }
  // nsx_core_c.c:17 kIndicatorTable (no closed form reproduces it exactly)
  static const int16_t kInd[17] = {0,    2017, 3809, 5227, 6258, 6963, 7424, 7718, 7901,
                                   8014, 8084, 8126, 8152, 8168, 8177, 8183, 8187};
  for (int i = 0; i < 17; ++i) t->indicator[i] = kInd[i];
  auto sums = [&](int from, int16_t* s1o, int16_t* s2o, int16_t* deto) {
    double s1 = 0, s2 = 0;
    for (int j = from; j < 129; ++j) {
      s1 += log2((double)j);
      s2 += log2((double)j) * log2((double)j);
    }
    *s1o = (int16_t)nsx_rnd(32.0 * s1);
    *s2o = (int16_t)nsx_rnd(4.0 * s2);
    if (deto) *deto = (int16_t)nsx_rnd((129 - from) * s2 - s1 * s1);
  };
  sums(5, &t->sum_log_idx5, &t->sum_sq_log_idx5, &t->det5);
  sums(65, &t->sum_log_idx65, &t->sum_sq_log_idx65, nullptr);
  // shared-memory images (nsx_layout.h)
  for (int v = 0; v < 2; ++v) {
    uint32_t* img = t->img[v];
    memset(img, 0, sizeof(t->img[v]));
    memcpy(img, v == 0 ? (const void*)t->win256 : (const void*)t->win128, v == 0 ? sizeof(t->win256) : sizeof(t->win128));
    memcpy(img + 128, t->tw, sizeof(t->tw));
    memcpy(img + 256, t->log_frac, sizeof(t->log_frac));
    nsx_fill_fft_twiddles(t->tw, v == 0 ? 256 : 128, reinterpret_cast<int32_t*>(img + kNsxImgFftTw));
  }
}

inline void nsx_set_mode(uint32_t* slab, int mode) {
  static const uint32_t od[4] = {256, 256, 282, 320};       // nsx_core.c:793-808
  static const uint32_t db[4] = {8192, 4096, 2048, 1475};
  if (mode < 0 || mode > 3) mode = 0;
  slab[kX_overdrive] = od[mode];
  slab[kX_denoiseBound] = db[mode];
  slab[kX_gainMap] = mode != 0;
  slab[kX_mode] = (uint32_t)mode;
}

inline void nsx_init_state(uint32_t* slab, uint32_t fs) {
  memset(slab, 0, sizeof(uint32_t) * (size_t)kNsxStateWords);
  int32_t* i = reinterpret_cast<int32_t*>(slab);
  const int32_t thr = fs == 8000 ? 131072 : 212644;
  i[kX_blockIndex] = -1;
  for (int s = 0; s < 3; ++s) i[kX_counter + s] = (200 * (s + 1)) / 3;
  i[kX_minNorm] = 15;
  i[kX_priorNonSpeech] = 8192;
  i[kX_thrLrt] = thr;
  i[kX_featLrt] = thr;
  i[kX_wLrt] = 6;
  i[kX_thrDiff] = 50;
  i[kX_featDiff] = 50;
  i[kX_thrFlat] = 20480;
  i[kX_featFlat] = 20480;
  i[kX_fs] = (int32_t)fs;
  i[kX_initFlag] = 1;
  for (int b = 0; b < 129; ++b) {
    uint32_t* a = slab + kNsxOffRecA + 4 * b;
    a[0] = 2048u | (2048u << 16);   // lq0 | lq1   (Q8 8.0)
    a[1] = 2048u | (153u << 16);    // lq2 | dens0 (Q9 0.3)
    a[2] = 153u | (153u << 16);     // dens1 | dens2
    a[3] = 0u | (16384u << 16);     // quantile | noiseSupFilter (Q14 1.0)
  }
  nsx_set_mode(slab, 0);
}

// Control words of a state slab that arrives from outside (WebRtcNsB200_ImportState): the kernel indexes the
// 1 / (counter + 1) table with the counters and shifts by the Q-domain words.
inline bool nsx_header_sane(const uint32_t* slab, uint32_t fs, int mode) {
  const int32_t* w = reinterpret_cast<const int32_t*>(slab);
  if (w[kX_blockIndex] < -1) return false;
  for (int s = 0; s < 3; ++s)
    if (w[kX_counter + s] < 0 || w[kX_counter + s] > 200) return false;
  if (w[kX_cntThresUpdate] < 0 || w[kX_cntThresUpdate] > 500) return false;
  const int q[4] = {w[kX_minNorm], w[kX_qNoise], w[kX_prevQNoise], w[kX_prevQMagn]};
  for (int k = 0; k < 4; ++k)
    if (q[k] < -64 || q[k] > 64) return false;
  return (uint32_t)w[kX_fs] == fs && w[kX_mode] == mode;
}

}  // namespace nsb200

#endif  // AUDIOSIGNALPROCESS_B200_NSX_HOST_INIT_H_
