// PCIe probe for the host-pointer batch path: DMA copies (1-D, 2-D with short rows, duplex) against
// zero-copy kernels that read/write pinned host memory directly, shaped like the NS kernel's PCM
// traffic (one warp per stream, 320-byte frames).   nvcc -O3 -arch=sm_100a -o tools/bin/pcie_probe tools/pcie_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// warp per stream, frame by frame: read 80 words (160 int16) from `in`, write to `out`; `work` spins
__global__ void zc_kernel(const uint32_t* in, uint32_t* out, int n_streams, int frames, size_t stride_w, int depth, int spin) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_streams) return;
  const uint32_t* src = in ? in + warp * stride_w : nullptr;
  uint32_t* dst = out ? out + warp * stride_w : nullptr;
  uint32_t acc = 0;
  uint32_t q[4][3];
  for (int d = 0; d < depth && d < frames; ++d)
    for (int u = 0; u < 3; ++u) q[d][u] = (src && lane + 32 * u < 80) ? __ldcs(src + d * 80 + lane + 32 * u) : 0u;
  for (int f = 0; f < frames; ++f) {
    uint32_t cur[3];
    const int slot = f % depth;
#pragma unroll
    for (int d = 0; d < 4; ++d) if (d == slot) for (int u = 0; u < 3; ++u) cur[u] = q[d][u];
    if (f + depth < frames) {
#pragma unroll
      for (int d = 0; d < 4; ++d) if (d == slot)
        for (int u = 0; u < 3; ++u) q[d][u] = (src && lane + 32 * u < 80) ? __ldcs(src + (f + depth) * 80 + lane + 32 * u) : 0u;
    }
    for (int s = 0; s < spin; ++s) acc = acc * 1664525u + cur[s % 3];
    if (dst) for (int u = 0; u < 3; ++u) if (lane + 32 * u < 80) __stcs(dst + f * 80 + lane + 32 * u, cur[u] + acc);
  }
}

#include <string.h>
#include <time.h>
#include <unistd.h>
// `pcie_probe dma [seconds]`: DMA copies only, each kind sustained for `seconds` (default 2) -- meant to be started
// on several GPUs at once (one process per GPU, CUDA_VISIBLE_DEVICES), to see what the host fabric gives each GPU
// when all of them copy: tools/pcie_probe_multi.sh.
static double now_s() { struct timespec t; clock_gettime(CLOCK_REALTIME, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char** argv) {
  const int S = 4096, F = 100, FL = 160;
  if (argc > 1 && strcmp(argv[1], "dma") == 0) {
    const double secs = argc > 2 ? atof(argv[2]) : 2.0;
    const size_t stride = (size_t)F * FL, bytes = S * stride * 2;
    int16_t *h_in, *h_out, *d_in, *d_out;
    CK(cudaMallocHost(&h_in, bytes)); CK(cudaMallocHost(&h_out, bytes));
    CK(cudaMalloc(&d_in, bytes)); CK(cudaMalloc(&d_out, bytes));
    memset(h_in, 1, bytes);
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    // all processes start their phases on the same wall-clock marks
    double t_mark = (double)((long)now_s() / 4 * 4 + 8);
    const int cf = 25;   // 25-frame chunks: rows of 8000 bytes at a pitch of 32000, the host-pointer path's copies
    for (int kind = 0; kind < 4; ++kind) {
      while (now_s() < t_mark) usleep(200);
      const double t0 = now_s();
      long reps = 0;
      while (now_s() - t0 < secs) {
        if (kind == 0) {
          CK(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s1));
          CK(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s2));
        } else if (kind == 1) {
          for (int f0 = 0; f0 < F; f0 += cf) {
            CK(cudaMemcpy2DAsync(d_in + (size_t)f0 * FL * S, (size_t)cf * FL * 2, h_in + f0 * FL, stride * 2, (size_t)cf * FL * 2, S, cudaMemcpyHostToDevice, s1));
            CK(cudaMemcpy2DAsync(h_out + f0 * FL, stride * 2, d_out + (size_t)f0 * FL * S, (size_t)cf * FL * 2, (size_t)cf * FL * 2, S, cudaMemcpyDeviceToHost, s2));
          }
        } else if (kind == 2) {
          CK(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s1));
        } else {
          CK(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s2));
        }
        CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2));
        ++reps;
      }
      const double dt = now_s() - t0;
      static const char* names[4] = {"duplex 1-D (per direction)", "duplex 2-D rows 8000 B pitch 32000 (per direction)", "H2D 1-D alone", "D2H 1-D alone"};
      printf("%-52s %6.1f GB/s\n", names[kind], reps * (double)bytes / dt * 1e-9);
      fflush(stdout);
      t_mark += secs + 2.0;
    }
    return 0;
  }
  const size_t stride = (size_t)F * FL, bytes = S * stride * 2;
  int16_t *h_in, *h_out, *d_in, *d_out;
  CK(cudaMallocHost(&h_in, bytes)); CK(cudaMallocHost(&h_out, bytes));
  CK(cudaMalloc(&d_in, bytes)); CK(cudaMalloc(&d_out, bytes));
  for (size_t i = 0; i < bytes / 2; ++i) h_in[i] = (int16_t)i;
  cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, auto fn, double gb) {
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0, s1)); fn();
      CK(cudaStreamSynchronize(s2)); CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf("%-46s %7.3f ms  %6.1f GB/s\n", name, best, gb / best * 1e-6);
  };
  const double GB = (double)bytes;
  timeit("H2D 1-D 131 MB", [&] { CK(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s1)); }, GB);
  timeit("D2H 1-D 131 MB", [&] { CK(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s1)); }, GB);
  timeit("duplex 1-D (per direction)", [&] { CK(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s1));
                                             CK(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s2)); }, GB);
  for (int cf : {5, 10, 13, 25, 50}) {
    char nm[96]; snprintf(nm, 96, "H2D 2-D rows of %d B (all chunks)", cf * FL * 2);
    timeit(nm, [&] { for (int f0 = 0; f0 < F; f0 += cf) { int nf = F - f0 < cf ? F - f0 : cf;
      CK(cudaMemcpy2DAsync(d_in + (size_t)f0 * FL * S, (size_t)nf * FL * 2, h_in + f0 * FL, stride * 2, (size_t)nf * FL * 2, S, cudaMemcpyHostToDevice, s1)); } }, GB);
    snprintf(nm, 96, "D2H 2-D rows of %d B (all chunks)", cf * FL * 2);
    timeit(nm, [&] { for (int f0 = 0; f0 < F; f0 += cf) { int nf = F - f0 < cf ? F - f0 : cf;
      CK(cudaMemcpy2DAsync(h_out + f0 * FL, stride * 2, d_out + (size_t)f0 * FL * S, (size_t)nf * FL * 2, (size_t)nf * FL * 2, S, cudaMemcpyDeviceToHost, s1)); } }, GB);
  }
  const int blocks = S / 2;
  for (int depth : {1, 2, 4}) for (int spin : {0, 2000}) {
    char nm[96];
    snprintf(nm, 96, "zero-copy read  depth %d spin %d", depth, spin);
    timeit(nm, [&] { zc_kernel<<<blocks, 64, 0, s1>>>((const uint32_t*)h_in, (uint32_t*)d_out, S, F, stride / 2, depth, spin); }, GB);
    snprintf(nm, 96, "zero-copy write depth %d spin %d", depth, spin);
    timeit(nm, [&] { zc_kernel<<<blocks, 64, 0, s1>>>((const uint32_t*)d_in, (uint32_t*)h_out, S, F, stride / 2, depth, spin); }, GB);
    snprintf(nm, 96, "zero-copy r+w   depth %d spin %d (per dir)", depth, spin);
    timeit(nm, [&] { zc_kernel<<<blocks, 64, 0, s1>>>((const uint32_t*)h_in, (uint32_t*)h_out, S, F, stride / 2, depth, spin); }, GB);
  }
  timeit("device-only kernel depth 2 spin 2000", [&] { zc_kernel<<<blocks, 64, 0, s1>>>((const uint32_t*)d_in, (uint32_t*)d_out, S, F, stride / 2, 2, 2000); }, GB);
  CK(cudaDeviceSynchronize());
  // verify zero-copy r+w result equals device path
  return 0;
}
