// GPU-box microbenchmark behind two design decisions of the float kernel:
//   * do packed fp32 instructions (FADD2 / FMUL2 / FFMA2, sm_100) issue as fast as the scalar ones, i.e. do they
//     halve the issue slots of per-bin arithmetic that has two independent slots to pair?
//   * how fast is the FP64 pipe (the tracker's logarithm and exponential run there)?
// Each kernel keeps 8 independent dependency chains per thread, 4 warps per scheduler (the float kernel's occupancy).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/microbench_pipes tools/microbench_pipes.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, int iters, float seed) {
  float a[8];
  float2 p[8];
  double d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + i + threadIdx.x;
    p[i] = make_float2(a[i], a[i] + 0.5f);
    d[i] = a[i];
  }
  const float c = seed * 0.999f;
  const float2 c2 = make_float2(c, c + 1e-3f);
  const double cd = c;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) a[i] = __fmaf_rn(a[i], c, c);
        if (MODE == 1) p[i] = __ffma2_rn(p[i], c2, c2);
        if (MODE == 2) a[i] = __fadd_rn(a[i], c);
        if (MODE == 3) p[i] = __fadd2_rn(p[i], c2);
        if (MODE == 4) d[i] = fma(d[i], cd, cd);
        if (MODE == 5) a[i] = __fmul_rn(a[i], c);
        if (MODE == 6) p[i] = __fmul2_rn(p[i], c2);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + p[i].x + p[i].y + (float)d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(const char* name, int elems_per_instr) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 4 * 128);
  const int iters = 4000;
  k<MODE><<<sms * 4, 128>>>(out, 10, 1.0001f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<sms * 4, 128>>>(out, iters, 1.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double warp_instr_per_sched = (double)iters * 64 * 4;                  // 4 warps per scheduler
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-8s %8.3f ms  %.2f cycles per warp instruction per scheduler  (%.1f elements/clk/SM)\n", name, ms,
         cycles / warp_instr_per_sched, warp_instr_per_sched * 4 * 32 * elems_per_instr / cycles);
  cudaFree(out);
  return ms;
}

int main() {
  run<0>("FFMA", 1);
  run<1>("FFMA2", 2);
  run<2>("FADD", 1);
  run<3>("FADD2", 2);
  run<5>("FMUL", 1);
  run<6>("FMUL2", 2);
  run<4>("DFMA", 1);
  return 0;
}
