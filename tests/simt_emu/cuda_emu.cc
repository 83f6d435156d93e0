// TEST TOOL ONLY: thread-per-CUDA-thread launcher for cuda_emu.h.
#include "cuda_emu.h"

#include <stdlib.h>

#include <vector>

namespace simt_emu {

thread_local ThreadCtx ctx;

namespace {
struct Start {
  ThreadCtx c;
  void (*fn)(void*);
  void* args;
};
void* Tramp(void* p) {
  Start* s = (Start*)p;
  ctx = s->c;
  s->fn(s->args);
  return NULL;
}
}  // namespace

void launch(void (*thread_main)(void*), void* args, int grid, int block) {
  const int nwarps = (block + 31) / 32;
  for (int b = 0; b < grid; ++b) {
    std::vector<Warp> warps(nwarps);
    Block blk;
    pthread_barrier_init(&blk.bar, NULL, block);
    for (int w = 0; w < nwarps; ++w) {
      const int n = (w == nwarps - 1) ? block - 32 * w : 32;
      pthread_barrier_init(&warps[w].bar, NULL, n);
    }
    std::vector<pthread_t> th(block);
    std::vector<Start> st(block);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 1 << 20);
    for (int t = 0; t < block; ++t) {
      ThreadCtx c;
      c.tid = make_uint3(t, 0, 0);
      c.bid = make_uint3(b, 0, 0);
      c.bdim = dim3(block, 1, 1);
      c.gdim = dim3(grid, 1, 1);
      c.warp = &warps[t / 32];
      c.block = &blk;
      c.parity = 0;
      st[t].c = c;
      st[t].fn = thread_main;
      st[t].args = args;
      if (pthread_create(&th[t], &attr, Tramp, &st[t]) != 0) abort();
    }
    for (int t = 0; t < block; ++t) pthread_join(th[t], NULL);
    pthread_attr_destroy(&attr);
    for (int w = 0; w < nwarps; ++w) pthread_barrier_destroy(&warps[w].bar);
    pthread_barrier_destroy(&blk.bar);
  }
}

}  // namespace simt_emu
