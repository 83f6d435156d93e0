// Tuning/diagnostic tool (GPU box): every packed-fp32 helper of ns_warp.cuh / nsf_kernel.cuh against plain scalar
// code on random operands, bit for bit.   nvcc -std=c++17 -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a \
//   -Iaudiosignalprocess_b200/csrc tools/packed_selftest.cu -o tools/bin/packed_selftest && tools/bin/packed_selftest
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ns_warp.cuh"
#include "nsf_kernel.cuh"
using namespace nsb200;

__device__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }
__device__ float rf(uint32_t& s, float lo, float hi) { return lo + (hi - lo) * (float)(rng(s) >> 8) * (1.f / 16777216.f); }
__device__ bool same(float a, float b) { return __float_as_uint(a) == __float_as_uint(b); }
__device__ bool same2(float2 a, float2 b) { return same(a.x, b.x) && same(a.y, b.y); }

__device__ float s_fdiv(float a, float b) { return fdiv(a, b); }

__global__ void k(unsigned long long* bad, int iters) {
  uint32_t s = 12345u + 977u * (blockIdx.x * blockDim.x + threadIdx.x);
  for (int it = 0; it < iters; ++it) {
    const float2 a = make_float2(rf(s, -3.f, 3.f), rf(s, -3.f, 3.f));
    const float2 b = make_float2(rf(s, -3.f, 3.f), rf(s, -3.f, 3.f));
    const float2 c = make_float2(rf(s, 0.1f, 900.f), rf(s, 0.1f, 900.f));
    const float2 w = make_float2(rf(s, -1.f, 1.f), rf(s, -1.f, 1.f));
    int t = 0;
    auto chk = [&](bool ok) { if (!ok) atomicAdd(bad + t, 1ull); ++t; };
    chk(same2(cadd(a, b), make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y))));                       // 0
    chk(same2(csub(a, b), make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y))));                       // 1
    chk(same2(cadd_i(a, b), make_float2(__fsub_rn(a.x, b.y), __fadd_rn(a.y, b.x))));                     // 2
    chk(same2(csub_i(a, b), make_float2(__fadd_rn(a.x, b.y), __fsub_rn(a.y, b.x))));                     // 3
    chk(same2(cmul(a, b), make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                                      __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)))));            // 4
    chk(same2(cmul_conj(a, b), make_float2(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                                           __fsub_rn(__fmul_rn(a.y, b.x), __fmul_rn(a.x, b.y)))));       // 5
    chk(same2(vmadd(a, b, c), make_float2(__fadd_rn(__fmul_rn(a.x, b.x), c.x), __fadd_rn(__fmul_rn(a.y, b.y), c.y))));   // 6
    chk(same2(vmmadd(a, b, c, w), make_float2(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(c.x, w.x)),
                                              __fadd_rn(__fmul_rn(a.y, b.y), __fmul_rn(c.y, w.y)))));    // 7
    chk(same2(vfdiv(a, c), make_float2(s_fdiv(a.x, c.x), s_fdiv(a.y, c.y))));                            // 8
    chk(same2(vsqrt_p1(c), make_float2(nsb_sqrtf_p1(c.x), nsb_sqrtf_p1(c.y))));                          // 9
    {
      float2 v[4] = {a, b, w, make_float2(c.x * 0.01f, c.y * 0.01f)};
      float2 r[4] = {v[0], v[1], v[2], v[3]};
      radix4<-1>(v);
      const float2 s02 = make_float2(__fadd_rn(r[0].x, r[2].x), __fadd_rn(r[0].y, r[2].y));
      const float2 d02 = make_float2(__fsub_rn(r[0].x, r[2].x), __fsub_rn(r[0].y, r[2].y));
      const float2 s13 = make_float2(__fadd_rn(r[1].x, r[3].x), __fadd_rn(r[1].y, r[3].y));
      const float2 d13 = make_float2(__fsub_rn(r[1].x, r[3].x), __fsub_rn(r[1].y, r[3].y));
      const float2 id13 = make_float2(d13.y, -d13.x);
      chk(same2(v[0], make_float2(__fadd_rn(s02.x, s13.x), __fadd_rn(s02.y, s13.y))));                   // 10
      chk(same2(v[2], make_float2(__fsub_rn(s02.x, s13.x), __fsub_rn(s02.y, s13.y))));                   // 11
      chk(same2(v[1], make_float2(__fadd_rn(d02.x, id13.x), __fadd_rn(d02.y, id13.y))));                 // 12
      chk(same2(v[3], make_float2(__fsub_rn(d02.x, id13.x), __fsub_rn(d02.y, id13.y))));                 // 13
    }
    t = 22;
    {   // the per-bin phases: the float2 instantiation against two float instantiations
      struct NoChain { __device__ void advance(int) {} } nc;
      const float c1[3] = {(float)(1 + (rng(s) % 200)), (float)(1 + (rng(s) % 200)), (float)(1 + (rng(s) % 200))};
      float rc1[3], cf[3];
      for (int q = 0; q < 3; ++q) { rc1[q] = frcp_nr(c1[q]); cf[q] = c1[q] - 1.f; }
      const float2 lm = make_float2(rf(s, 0.f, 12.f), rf(s, 0.f, 12.f));
      float2 lq[3], dn[3];
      float lqa[3], lqb[3], dna[3], dnb[3];
      for (int q = 0; q < 3; ++q) {
        // a third of the operands close enough to lm for the density update to fire
        lq[q] = (rng(s) % 3) ? make_float2(rf(s, 0.f, 12.f), rf(s, 0.f, 12.f)) : make_float2(lm.x + rf(s, -0.02f, 0.02f), lm.y + rf(s, -0.02f, 0.02f));
        dn[q] = make_float2(rf(s, 0.01f, 60.f), rf(s, 0.01f, 60.f));
        lqa[q] = lq[q].x; lqb[q] = lq[q].y; dna[q] = dn[q].x; dnb[q] = dn[q].y;
      }
      nsf_tracker_update(lm, lq, dn, c1, rc1, cf, nc, 0);
      nsf_tracker_update(lm.x, lqa, dna, c1, rc1, cf, nc, 0);
      nsf_tracker_update(lm.y, lqb, dnb, c1, rc1, cf, nc, 0);
      bool ok = true;
      for (int q = 0; q < 3; ++q) ok = ok && same2(lq[q], make_float2(lqa[q], lqb[q])) && same2(dn[q], make_float2(dna[q], dnb[q]));
      chk(ok);   // 22
      const float2 magn = make_float2(rf(s, 1.f, 5000.f), rf(s, 1.f, 5000.f)), noise = make_float2(rf(s, 1.f, 5000.f), rf(s, 1.f, 5000.f));
      const float2 noisePrev = make_float2(rf(s, 1.f, 5000.f), rf(s, 1.f, 5000.f)), magnPrev = make_float2(rf(s, 1.f, 5000.f), rf(s, 1.f, 5000.f));
      const float2 smooth = make_float2(rf(s, 0.09f, 1.f), rf(s, 0.09f, 1.f)), pause = make_float2(rf(s, 1.f, 5000.f), rf(s, 1.f, 5000.f));
      const float avgM = rf(s, 1.f, 5000.f), avgP = rf(s, 1.f, 5000.f);
      float2 pe, ll = make_float2(rf(s, -2.f, 30.f), rf(s, -2.f, 30.f)), t0, t1, t2;
      float pea, lla = ll.x, t0a, t1a, t2a, peb, llb = ll.y, t0b, t1b, t2b;
      nsf_snr_lrt(magn, noise, noisePrev, magnPrev, smooth, pause, avgM, avgP, pe, ll, t0, t1, t2);
      nsf_snr_lrt(magn.x, noise.x, noisePrev.x, magnPrev.x, smooth.x, pause.x, avgM, avgP, pea, lla, t0a, t1a, t2a);
      nsf_snr_lrt(magn.y, noise.y, noisePrev.y, magnPrev.y, smooth.y, pause.y, avgM, avgP, peb, llb, t0b, t1b, t2b);
      chk(same2(pe, make_float2(pea, peb)) && same2(ll, make_float2(lla, llb)) && same2(t0, make_float2(t0a, t0b)) &&
          same2(t1, make_float2(t1a, t1b)) && same2(t2, make_float2(t2a, t2b)));   // 23
      { const int keep = t; t = 27;
        chk(same2(pe, make_float2(pea, peb)));
        chk(same2(ll, make_float2(lla, llb)));
        chk(same2(t0, make_float2(t0a, t0b)));
        chk(same2(t1, make_float2(t1a, t1b)));
        chk(same2(t2, make_float2(t2a, t2b)));
        chk(same2(vlog_rn(magn), make_float2(nsb_log_rn(magn.x), nsb_log_rn(magn.y))));
        chk(same2(vexp_rn(a), make_float2(nsb_exp_rn(a.x), nsb_exp_rn(a.y))));
        const float2 e = vmmadd(a, a, b, b);
        const float2 m = vsqrt_p1(e);
        chk(same2(vlog_rn(m), make_float2(nsb_log_rn(nsb_sqrtf_p1(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(b.x, b.x)))),
                                          nsb_log_rn(nsb_sqrtf_p1(__fadd_rn(__fmul_rn(a.y, a.y), __fmul_rn(b.y, b.y)))))));
        t = keep; }
      const float2 ps = make_float2(rf(s, 0.f, 1.f), rf(s, 0.f, 1.f));
      const bool2v ph = {(rng(s) & 1) != 0, (rng(s) & 1) != 0};
      float2 mp = pause, nz;
      float mpa = pause.x, mpb = pause.y, nza, nzb;
      nsf_noise_update(ps, ph, magn, noisePrev, mp, nz);
      nsf_noise_update(ps.x, ph.x, magn.x, noisePrev.x, mpa, nza);
      nsf_noise_update(ps.y, ph.y, magn.y, noisePrev.y, mpb, nzb);
      chk(same2(mp, make_float2(mpa, mpb)) && same2(nz, make_float2(nza, nzb)));   // 24
      const float2 gw = nsf_wiener_gain(magn, noise, pe, 1.25f, 0.09f);
      chk(same2(gw, make_float2(nsf_wiener_gain(magn.x, noise.x, pe.x, 1.25f, 0.09f), nsf_wiener_gain(magn.y, noise.y, pe.y, 1.25f, 0.09f))));   // 25
      const float2 X = nsf_real_split(a, b, w);
      const float xd = __fsub_rn(a.x, b.x), xs = __fadd_rn(a.y, b.y);
      chk(same2(X, make_float2(__fsub_rn(a.x, __fsub_rn(__fmul_rn(w.x, xd), __fmul_rn(w.y, xs))),
                               __fsub_rn(a.y, __fadd_rn(__fmul_rn(w.x, xs), __fmul_rn(w.y, xd))))));   // 26
    }
    t = 14;
    for (int diag = 0; diag < 2; ++diag) {
      float2 v[4] = {a, b, w, make_float2(c.x * 0.01f, c.y * 0.01f)};
      const float2 w1 = diag ? make_float2(0.70710678f, 0.f) : w, w2 = make_float2(b.x * 0.3f, a.y * 0.3f),
                   w3 = diag ? make_float2(-0.70710678f, 0.f) : make_float2(w.y, -w.x);
      // scalar form (the previous implementation)
      const float x0r = __fadd_rn(v[0].x, v[1].x), x0i = __fadd_rn(v[0].y, v[1].y);
      const float x1r = __fsub_rn(v[0].x, v[1].x), x1i = __fsub_rn(v[0].y, v[1].y);
      const float x2r = __fadd_rn(v[2].x, v[3].x), x2i = __fadd_rn(v[2].y, v[3].y);
      const float x3r = __fsub_rn(v[2].x, v[3].x), x3i = __fsub_rn(v[2].y, v[3].y);
      const float dr = __fsub_rn(x0r, x2r), di = __fsub_rn(x0i, x2i);
      const float yr = __fsub_rn(x1r, x3i), yi = __fadd_rn(x1i, x3r);
      const float zr = __fadd_rn(x1r, x3i), zi = __fsub_rn(x1i, x3r);
      const float2 e0 = make_float2(__fadd_rn(x0r, x2r), __fadd_rn(x0i, x2i));
      const float2 e2 = make_float2(__fsub_rn(__fmul_rn(w2.x, dr), __fmul_rn(w2.y, di)), __fadd_rn(__fmul_rn(w2.x, di), __fmul_rn(w2.y, dr)));
      const float tyr = diag ? yr : 0.f, tyi = diag ? yi : 0.f, tzr = diag ? zr : 0.f, tzi = diag ? zi : 0.f;
      const float2 e1 = make_float2(__fsub_rn(__fmul_rn(w1.x, __fsub_rn(yr, tyi)), __fmul_rn(w1.y, yi)),
                                    __fadd_rn(__fmul_rn(w1.x, __fadd_rn(yi, tyr)), __fmul_rn(w1.y, yr)));
      const float2 e3 = make_float2(__fsub_rn(__fmul_rn(w3.x, __fadd_rn(zr, tzi)), __fmul_rn(w3.y, zi)),
                                    __fadd_rn(__fmul_rn(w3.x, __fsub_rn(zi, tzr)), __fmul_rn(w3.y, zr)));
      ooura_bfly(v, w1, w2, w3, diag != 0);
      chk(same2(v[0], e0));   // 14, 18
      chk(same2(v[1], e1));   // 15, 19
      chk(same2(v[2], e2));   // 16, 20
      chk(same2(v[3], e3));   // 17, 21
    }
  }
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 64 * 8);
  cudaMemset(d, 0, 64 * 8);
  k<<<148, 256>>>(d, 400);
  unsigned long long h[64];
  cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
  const char* names[] = {"cadd", "csub", "cadd_i", "csub_i", "cmul", "cmul_conj", "vmadd", "vmmadd", "vfdiv", "vsqrt_p1",
                         "radix4[0]", "radix4[2]", "radix4[1]", "radix4[3]", "bfly a0", "bfly a1", "bfly a2", "bfly a3",
                         "bfly diag a0", "bfly diag a1", "bfly diag a2", "bfly diag a3",
                         "tracker", "snr_lrt", "noise_update", "wiener_gain", "real_split",
                         "snr: prevEst", "snr: logLrt", "snr: dm*dp", "snr: dp*dp", "snr: dm*dm", "vlog_rn", "vexp_rn", "lmagn"};
  int rc = 0;
  for (int i = 0; i < 35; ++i) { printf("%-14s mismatches %llu of %d\n", names[i], h[i], 148 * 256 * 400); rc |= h[i] != 0; }
  return rc;
}
