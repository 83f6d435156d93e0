// PLACEHOLDER until the fixed-point kernel lands (next milestone): lets the
// host layer compile; every NSx launch fails loudly.
#ifndef AUDIOSIGNALPROCESS_B200_NSX_KERNEL_CUH_
#define AUDIOSIGNALPROCESS_B200_NSX_KERNEL_CUH_
#include <stdint.h>
namespace nsb200 {
enum : int { kNsxHdrWords = 32, kNsxStateWords = 32 };
struct NsxTables { int unused; };
struct NsxLaunch {
  uint32_t* state; const int* slots; const NsxTables* tables; const void* in; void* out;
  long long in_stream_stride, in_frame_stride, in_band_stride;
  long long out_stream_stride, out_frame_stride, out_band_stride;
  int n_streams, frames;
};
constexpr int kNsxWarpsPerCta = 4;
constexpr int kNsxCtaTableWords = 0;
constexpr int kNsxWarpWords = 0;
template <int ANA, int NB>
__global__ void nsx_process_kernel(const NsxLaunch p) { __trap(); }
}  // namespace nsb200
#endif
