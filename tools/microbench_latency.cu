// Tuning tool (GPU box): dependent-chain latency of the instruction kinds the float kernel's serial sections are
// made of, one warp alone on an SM, cycles per dependent instruction.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench_latency.cu -o tools/bin/microbench_latency
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, long long* cyc, float seed, double dseed) {
  __shared__ float sm[64];
  const int lane = threadIdx.x;
  sm[lane] = (float)((lane + 1) & 31);
  sm[lane + 32] = (float)lane;
  __syncwarp();
  constexpr int N = 2048;
  long long t0, t1;
  float a = seed;
  double d = dseed;
  int idx = lane;
  // FFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = fmaf(a, 1.0000001f, 0.5f);
  t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
  // FADD2 (packed)
  float2 p = make_float2(a, a + 1.f);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) p = __fadd2_rn(p, make_float2(0.5f, 0.25f));
  t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
  a += p.x + p.y;
  // DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) d = fma(d, 1.00000001, 0.5);
  t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;
  // MUFU.RCP
  float r = a * 1e-9f + 1.5f;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(r));
  t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
  // LDS (pointer chase)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) idx = (int)sm[idx];
  t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
  // SHFL
  float s = r;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) s = __shfl_xor_sync(0xffffffffu, s, 1);
  t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
  // F2F.F64.F32 + F2F.F32.F64 round trip
  float c = s + 1.f;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double q = (double)c; asm volatile("" : "+d"(q)); c = (float)q; asm volatile("" : "+f"(c)); }
  t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0;
  // REDUX
  int v = idx + lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) v = __reduce_max_sync(0xffffffffu, v) + lane;
  t1 = clock64(); if (lane == 0) cyc[7] = t1 - t0;
  out[lane] = a + (float)d + r + s + c + (float)idx + (float)v;
}
int main() {
  float* o; long long* c;
  cudaMalloc(&o, 128); cudaMalloc(&c, 64);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(o, c, 1.f, 1.0);
  long long h[8];
  cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
  const char* n[] = {"FFMA", "FADD2", "DFMA", "MUFU.RCP", "LDS + F2I (index chase)", "SHFL", "F2F f32->f64->f32 (pair)", "REDUX + IADD"};
  for (int i = 0; i < 8; ++i)
    if (i != 3 && i != 5)   // (the compiler folds these two chains away: not measured)
      printf("%-26s %.1f cycles per dependent step\n", n[i], (double)h[i] / 2048.0);
  return 0;
}
