// TEST TOOL ONLY.  A minimal SIMT emulator: lets the host compiler build the
// kernel sources under audiosignalprocess_b200/csrc/*.cuh unchanged and run
// them with one CPU thread per CUDA thread, so that kernel logic can be checked
// against the oracle in the CPU-only test suite (`-m "not gpu"`) and debugged
// without a GPU.  It is never linked into the product library and the product
// has no path that reaches it.
//
// Supported subset: 1-D grids/blocks, dynamic shared memory (one block runs at
// a time), full-warp shuffles / __syncwarp in warp-uniform control flow,
// __syncthreads, int atomicAdd, libm float math.
#ifndef TESTS_SIMT_EMU_CUDA_EMU_H_
#define TESTS_SIMT_EMU_CUDA_EMU_H_

#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <vector_functions.h>
#include <vector_types.h>

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __launch_bounds__
#undef __shared__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__
#define NSB_DEV static inline
#define NSB_DEVM inline

namespace simt_emu {

struct Warp {
  pthread_barrier_t bar;
  uint32_t buf[2][32];
};
struct Block {
  pthread_barrier_t bar;
};
struct ThreadCtx {
  uint3 tid, bid;
  dim3 bdim, gdim;
  Warp* warp;
  Block* block;
  int parity;
};
extern thread_local ThreadCtx ctx;

inline uint32_t xchg(uint32_t v, int src_lane) {
  Warp* w = ctx.warp;
  const int par = (ctx.parity ^= 1);
  w->buf[par][ctx.tid.x & 31u] = v;
  pthread_barrier_wait(&w->bar);
  return w->buf[par][src_lane & 31];
}
template <typename T>
inline T shfl(T v, int src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  uint32_t u;
  memcpy(&u, &v, 4);
  u = xchg(u, src_lane);
  memcpy(&v, &u, 4);
  return v;
}

// Runs kernel(args) over grid x block threads, one block at a time.
void launch(void (*thread_main)(void*), void* args, int grid, int block);

}  // namespace simt_emu

#define threadIdx (simt_emu::ctx.tid)
#define blockIdx (simt_emu::ctx.bid)
#define blockDim (simt_emu::ctx.bdim)
#define gridDim (simt_emu::ctx.gdim)

template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return simt_emu::shfl(v, src); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) {
  return simt_emu::shfl(v, (int)(simt_emu::ctx.tid.x & 31u) ^ m);
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, int d) {
  const int lane = (int)(simt_emu::ctx.tid.x & 31u);
  T o = simt_emu::shfl(v, lane - d >= 0 ? lane - d : lane);
  return lane - d >= 0 ? o : v;
}
template <typename T> inline T __shfl_down_sync(unsigned, T v, int d) {
  const int lane = (int)(simt_emu::ctx.tid.x & 31u);
  T o = simt_emu::shfl(v, lane + d < 32 ? lane + d : lane);
  return lane + d < 32 ? o : v;
}
inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned r = 0;
  for (int l = 0; l < 32; ++l) r |= (simt_emu::shfl(pred ? 1u : 0u, l) & 1u) << l;
  return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&simt_emu::ctx.warp->bar); }
inline void __syncthreads() { pthread_barrier_wait(&simt_emu::ctx.block->bar); }
inline void __threadfence_block() { __sync_synchronize(); }
inline void __threadfence() { __sync_synchronize(); }
inline int atomicAdd(int* p, int v) { return __sync_fetch_and_add(p, v); }
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}

#endif  // TESTS_SIMT_EMU_CUDA_EMU_H_
