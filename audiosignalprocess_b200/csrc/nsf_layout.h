// Float-NS device data layout shared by the kernel (nsf_kernel.cuh), the C-ABI
// host layer and the test-only emulator: per-stream state slab, cold histogram
// slab, constant tables, launch descriptor.  Plain C++ (needs only float2).
#ifndef AUDIOSIGNALPROCESS_B200_NSF_LAYOUT_H_
#define AUDIOSIGNALPROCESS_B200_NSF_LAYOUT_H_

#include <stdint.h>
#include <vector_types.h>

namespace nsb200 {

// ---- state slab layout (32-bit words) --------------------------------------
enum : int {
  kH_blockInd = 0,       // int   ns_core.h:81
  kH_updates = 1,        // int   :70
  kH_counter = 2,        // int[3] :69
  kH_modelUpd0 = 5,      // int   modelUpdatePars[0] :82
  kH_modelUpd3 = 6,      // int   modelUpdatePars[3]
  kH_splitValid = 7,     // int   1: the split-mode arrays below (kNsfOffSplit) are current
  kH_priorPars = 8,      // float[7] priorModelPars :84
  kH_priorSpeechProb = 15,
  kH_feat = 16,          // float[7] featureData :93
  kH_white = 23,         // whiteNoiseLevel :98
  kH_pinkNum = 24,       // pinkNoiseNumerator :100
  kH_pinkExp = 25,       // pinkNoiseExp :101
  kH_overdrive = 26,     // :73
  kH_denoiseBound = 27,  // :74
  kH_gainmap = 28,       // int :75
  kH_mode = 29,          // int aggrMode
  kH_fs = 30,            // int
  kH_initFlag = 31,      // int
  kNsfHdrWords = 32,

  kNsfOffXHist = 32,                 // 96 floats (48 @ 8 kHz): last samples of dataBuf
  kNsfOffSynt = kNsfOffXHist + 96,   // 96 floats: head of syntBuf
  kNsfOffHb = kNsfOffSynt + 96,      // 2 x 96 floats: tails of dataBufHB[0..1]
  kNsfOffInitMagn = kNsfOffHb + 192, // 132 floats: initMagnEst
  kNsfOffBins = kNsfOffInitMagn + 132,  // 129 x 12 floats
  kNsfBinRec = 12,
  kNsfFusedWordsRaw = kNsfOffBins + 129 * kNsfBinRec,
  // Split mode only (Analyze and Process fed different signals): what the
  // fused kernel never needs because it is implied when both see one frame.
  kNsfOffSplit = (kNsfFusedWordsRaw + 31) / 32 * 32,
  kNsfOffPHist = kNsfOffSplit,       // 96 floats: last samples of dataBuf (kNsfOffXHist is analyzeBuf then)
  kNsfOffAux = kNsfOffPHist + 96,    // 4 x 132 floats: noise | speechProb | parametricNoise | magnPrevProcess
  kNsfAuxStride = 132,               //   (the record's magnPrev is magnPrevAnalyze then; ns_core.h:64-66,105-107)
  kNsfStateWordsRaw = kNsfOffAux + 4 * kNsfAuxStride,
  kNsfStateWords = (kNsfStateWordsRaw + 31) / 32 * 32,  // 128-byte aligned slabs
  kNsfHistWords = 3008,  // 3 x 1000 int32, padded to 128 B
};
// per-bin record fields
enum : int {
  kB_lq0 = 0, kB_lq1, kB_lq2, kB_dens0, kB_dens1, kB_dens2, kB_quantile, kB_smooth,
  kB_noisePrev, kB_magnPrev, kB_logLrt, kB_magnAvgPause
};

// Read-only per-device tables.  img[v] (v = 0: 256-point analysis, 1: 128-point) is the image of
// the kernel's per-CTA table block exactly as it sits in shared memory -- window | twiddles |
// log(i) | pad | twiddles of the inverse FFT's passes 1-2 regrouped per (factor, lane) (ns_warp.cuh
// fft_fill_tw12) | real-input split and twiddles of the forward FFT, which follows the reference's
// rounding order (ns_warp.cuh ooura_fwd) -- so that a CTA fetches it with one TMA bulk copy.
enum : int {
  kNsfImgWin = 0, kNsfImgTw = 256, kNsfImgLogi = 768, kNsfImgTw12 = 912,
  kNsfImgSplit = 1152,   // 129 float2: real-input split of the forward FFT per bin (nsf_host_init.h)
  kNsfImgOtw = 1412,     // 122 float2: forward-FFT twiddles per (pass, lane) (ns_warp.cuh ooura_fwd)
  kNsfTableImgWords = 1656,
};
struct NsfTables {
  alignas(16) float img[2][kNsfTableImgWords];
  float win256[256];  // kBlocks160w256  windows_private.h:94
  float win128[128];  // kBlocks80w128   windows_private.h:64
  float2 tw[256];     // e^{+2 pi i t/256}
  float logi[132];    // (float)log((float)i), i >= 1
  float sum_log_i[2];     // [0]: magnLen 129, [1]: magnLen 65   (ns_core.c:1093-1095)
  float sum_log_i_sq[2];
  double logk_d[132];     // log(k) in double, k >= 1: the start-up frames' pow(k, exp) = e^(exp log k)
};

struct NsfLaunch {
  float* state;            // slab base
  int* hist;               // cold histogram slab base
  const int* slots;        // [n_streams] slab index per batch entry; NULL: entry i sits in slot slot_base + i
  int slot_base = 0;
  int prefetch_ahead = 0;  // != 0: a warp pulls the state of batch entry (its own + |prefetch_ahead|) into L2; < 0: modulo the batch
  const NsfTables* tables;
  const void* in;          // int16 or float samples
  void* out;
  const void* ana_in;      // split mode: band-0 frames for Analyze (same sample type as `in`)
  long long ana_stream_stride, ana_frame_stride;
  long long in_stream_stride, in_frame_stride, in_band_stride;     // in elements
  long long out_stream_stride, out_frame_stride, out_band_stride;
  int n_streams;
  int frames;
  int phase = 0;           // split kernels: 1 = Analyze half only, 2 = Process half only, 0 = both (nsf_kernel.cuh)
};

}  // namespace nsb200

#endif  // AUDIOSIGNALPROCESS_B200_NSF_LAYOUT_H_
