// PLACEHOLDER until the QMF / resampler band-split kernels land.
#ifndef AUDIOSIGNALPROCESS_B200_BAND_KERNELS_CUH_
#define AUDIOSIGNALPROCESS_B200_BAND_KERNELS_CUH_
#include <cuda_runtime.h>
#include <stdint.h>
namespace nsb200 {
enum : int { kBandStateWords = 32 };
struct BandLaunch {
  int32_t* state; const int* slots; int16_t* full; long long full_stride; int16_t* bands;
  long long bands_stride; int n_streams, frames;
};
inline int LaunchBandSplit(int, const BandLaunch&, cudaStream_t, uint64_t*) { return -1; }
inline int LaunchBandMerge(int, const BandLaunch&, cudaStream_t, uint64_t*) { return -1; }
}  // namespace nsb200
#endif
