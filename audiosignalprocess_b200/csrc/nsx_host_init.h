// PLACEHOLDER (see nsx_kernel.cuh).
#ifndef AUDIOSIGNALPROCESS_B200_NSX_HOST_INIT_H_
#define AUDIOSIGNALPROCESS_B200_NSX_HOST_INIT_H_
#include <stdint.h>
namespace nsb200 {
template <typename T> inline void nsx_fill_tables(T*) {}
inline void nsx_init_state(uint32_t*, uint32_t) {}
inline void nsx_set_mode(uint32_t*, int) {}
}  // namespace nsb200
#endif
