// Fixed-point NSx device data layout shared by the kernel (nsx_kernel.cuh), the
// C-ABI host layer and the test-only emulator.  Plain C++.
//
// Per-stream slab (32-bit words): header scalars | analysis history | synthesis
// overlap | HB delay lines | initMagnEst | per-bin record A (8 x int16) |
// per-bin record B (4 x 32 bit) | cold: 3 x 1000 histogram counters.
// Field names follow NoiseSuppressionFixedC, ns/nsx_core.h:22-110.
#ifndef AUDIOSIGNALPROCESS_B200_NSX_LAYOUT_H_
#define AUDIOSIGNALPROCESS_B200_NSX_LAYOUT_H_

#include <stdint.h>

namespace nsb200 {

enum : int {
  kX_blockIndex = 0,      // nsx_core.h:85
  kX_counter = 1,         // [3] noiseEstCounter :34
  kX_minNorm = 4,         // :75
  kX_priorNonSpeech = 5,  // :83 (Q14)
  kX_cntThresUpdate = 6,  // :88
  kX_qNoise = 7,          // :98
  kX_prevQNoise = 8,      // :99
  kX_prevQMagn = 9,       // :100
  kX_featLrt = 10,        // featureLogLrt :49
  kX_thrLrt = 11,         // thresholdLogLrt :50
  kX_wLrt = 12, kX_wDiff = 13, kX_wFlat = 14,  // weights :51,55,59
  kX_featDiff = 15, kX_thrDiff = 16,           // :53,54
  kX_featFlat = 17, kX_thrFlat = 18,           // :57,58
  kX_curAvgEnergy = 19,   // curAvgMagnEnergy :65
  kX_timeAvgEnergy = 20,  // :66
  kX_timeAvgEnergyTmp = 21,  // :67
  kX_whiteLevel = 22,     // :69
  kX_pinkNum = 23,        // :73
  kX_pinkExp = 24,        // :74
  kX_overdrive = 25,      // :29 (Q8)
  kX_denoiseBound = 26,   // :30 (Q14)
  kX_gainMap = 27,
  kX_mode = 28,
  kX_fs = 29,
  kX_initFlag = 30,
  kNsxHdrWords = 32,

  kNsxOffAna = 32,                       // 96 int16: tail of analysisBuffer
  kNsxOffSyn = kNsxOffAna + 48,          // 96 int16: head of synthesisBuffer
  kNsxOffHb = kNsxOffSyn + 48,           // 2 x 96 int16: tails of dataBufHBFX
  kNsxOffInitMagn = kNsxOffHb + 96,      // 132 x uint32 initMagnEst
  kNsxOffRecA = kNsxOffInitMagn + 132,   // 129 x {lq0|lq1, lq2|dens0, dens1|dens2, quantile|filter}
  kNsxOffRecB = kNsxOffRecA + 129 * 4,   // 129 x {logLrtTimeAvgW32, avgMagnPause, prevNoiseU32, prevMagnU16}
  kNsxHotWords = kNsxOffRecB + 129 * 4,
  kNsxOffHist = (kNsxHotWords + 31) / 32 * 32,  // cold: histLrt | histSpecFlat | histSpecDiff (int32 counters)
  kNsxStateWords = (kNsxOffHist + 3000 + 31) / 32 * 32,
};

// Read-only tables (built by formula in nsx_host_init.h; the literal reference
// tables they reproduce are cited there).
// img[v] (v = 0: 256-point analysis, 1: 128-point) is the image of the kernel's per-CTA table block as
// it sits in shared memory (window | twiddles, packed | log2 fraction table | FFT twiddles as int2, regrouped
// per stage: ns_fixed.cuh), fetched with one TMA bulk copy.
enum : int { kNsxImgFftTw = 128 + 128 + 128, kNsxTableImgWords = kNsxImgFftTw + 2 * 252 };
struct NsxTables {
  alignas(16) uint32_t img[2][kNsxTableImgWords];
  int16_t win256[256];
  int16_t win128[128];
  uint32_t tw[128];         // (cos, sin)(2 pi t / 256) packed int16 pairs from kSinTable1024
  int16_t log_frac[256];
  int16_t counter_div[202];
  int16_t log_tab[10];
  int16_t log_idx[130];
  int16_t factor1[258];
  int16_t factor2[3][258];
  int16_t indicator[18];
  int16_t sum_log_idx5, sum_sq_log_idx5, det5;      // table entries at kStartBand = 5
  int16_t sum_log_idx65, sum_sq_log_idx65, pad_;
};

struct NsxLaunch {
  uint32_t* state;
  const int* slots;        // [n_streams] slab index per batch entry; NULL: entry i sits in slot slot_base + i
  int slot_base = 0;
  int prefetch_ahead = 0;  // > 0: a warp pulls the state of batch entry (its own + prefetch_ahead) into L2
  const NsxTables* tables;
  const void* in;   // int16 samples
  void* out;
  long long in_stream_stride, in_frame_stride, in_band_stride;
  long long out_stream_stride, out_frame_stride, out_band_stride;
  int n_streams;
  int frames;
};

}  // namespace nsb200

#endif  // AUDIOSIGNALPROCESS_B200_NSX_LAYOUT_H_
