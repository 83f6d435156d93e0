// Device-side fixed-point primitives of the NSx kernel: the integer helpers of
// WebRTC's signal processing library (common_audio/signal_processing/, "SPL")
// restated for one GPU thread, and the warp-wide int16 complex FFT pair that
// replaces WebRtcSpl_RealForwardFFT / RealInverseFFT (SPL/real_fft.c:47,74 ->
// complex_bit_reverse.c:49, complex_fft.c:29,160).  Everything here is
// bit-exact by construction: same operand widths, same truncating casts, same
// rounding constants, same butterfly order.
#ifndef AUDIOSIGNALPROCESS_B200_NS_FIXED_CUH_
#define AUDIOSIGNALPROCESS_B200_NS_FIXED_CUH_

#include "ns_warp.cuh"

namespace nsb200 {

// ---- scalar helpers (SPL/include/spl_inl.h, signal_processing_library.h) ----
NSB_DEV int fx_sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }  // spl_inl.h:27
NSB_DEV int fx_s16(int v) { return (int)(int16_t)v; }                                   // (int16_t) cast
NSB_DEV int fx_norm_u32(unsigned a) { return a == 0u ? 0 : __clz((int)a); }             // spl_inl.h:126
NSB_DEV int fx_norm_w32(int a) {                                                        // spl_inl.h:103
  if (a == 0) return 0;
  if (a < 0) a = ~a;
  return a == 0 ? 31 : __clz(a) - 1;
}
NSB_DEV int fx_norm_w16(int a16) {                                                      // spl_inl.h:144
  int v = fx_s16(a16);
  if (v == 0) return 0;
  if (v < 0) v = ~v;
  return v == 0 ? 15 : __clz(v) - 17;
}
NSB_DEV int fx_shift_w32(int x, int c) {  // WEBRTC_SPL_SHIFT_W32
  return c >= 0 ? (int)((unsigned)x << c) : (x >> (-c));
}
NSB_DEV int fx_shl(int x, int c) { return (int)((unsigned)x << c); }
NSB_DEV int fx_mul_rsft_round(int a16, int b16, int c) { return (a16 * b16 + (1 << (c - 1))) >> c; }

// floor(sqrt(v)) for the int32 image of v; 0 when that image is negative,
// which is what the restoring iteration of SPL/spl_sqrt_floor.c:55 returns.
NSB_DEV unsigned fx_sqrt_floor(unsigned v) {
  if ((int)v <= 0) return 0u;
  unsigned r = (unsigned)sqrtf((float)v);
  while (r * r > v) --r;
  while ((r + 1u) * (r + 1u) <= v) ++r;
  return r;
}

// ---- warp-wide int16 complex FFT -------------------------------------------
// N = 256 (8 stages) or 128 (7 stages) complex points, LANES = N/8 active lanes,
// 8 points per lane as packed words (re in the low half, im in the high half).
// Radix-2 decimation in time in exactly the reference's stage order; three
// register layouts, two conflict-free shared-memory transposes:
//   A: position p = 8*lane + r                      stages 0,1,2
//   B: p = (lane&7) + 8*r + 64*(lane>>3)            stages 3,4,5
//   C: p = lane + LANES*r                           stages 6(,7)
// tw[t] = (cos, sin)(2*pi*t/256) = (kSinTable1024[4t+256], kSinTable1024[4t]) packed.
NSB_DEV int fx_swz(int p) { return p ^ ((p >> 3) & 7) ^ (((p >> 6) & 3) << 3); }

NSB_DEV uint32_t fx_pack(int re, int im) { return ((uint32_t)re & 0xffffu) | ((uint32_t)im << 16); }
NSB_DEV int fx_lo(uint32_t w) { return (int)(int16_t)(w & 0xffffu); }
NSB_DEV int fx_hi(uint32_t w) { return (int)(int16_t)(w >> 16); }

// One butterfly of complex_fft.c mode 1. INV=false: forward (wi = -sin, fixed
// shift 1); INV=true: inverse (wi = +sin, `shift` in 0..2, round2 = 8192 << shift).
template <bool INV>
NSB_DEV void fx_butterfly(uint32_t& lo_w, uint32_t& hi_w, uint32_t tw, int shift) {
  const int wr = fx_lo(tw), ws = fx_hi(tw);
  const int xr = fx_lo(hi_w), xi = fx_hi(hi_w);
  int tr, ti;
  if (INV) {
    tr = (wr * xr - ws * xi + 1) >> 1;
    ti = (wr * xi + ws * xr + 1) >> 1;
  } else {
    tr = (wr * xr + ws * xi + 1) >> 1;
    ti = (wr * xi - ws * xr + 1) >> 1;
  }
  const int qr = fx_lo(lo_w) * 16384, qi = fx_hi(lo_w) * 16384;
  const int rnd = INV ? (8192 << shift) : 16384;
  const int sh = INV ? (shift + 14) : 15;
  hi_w = fx_pack((qr - tr + rnd) >> sh, (qi - ti + rnd) >> sh);
  lo_w = fx_pack((qr + tr + rnd) >> sh, (qi + ti + rnd) >> sh);
}

// max |int16| over the 8 packed words of this lane (no 32768 clamp yet)
NSB_DEV int fx_lane_max_abs(const uint32_t (&v)[8]) {
  int m = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int a = fx_lo(v[r]), b = fx_hi(v[r]);
    a = a < 0 ? -a : a;
    b = b < 0 ? -b : b;
    m = a > m ? a : m;
    m = b > m ? b : m;
  }
  return m;
}

// Runs the stages whose butterfly span lies inside the lane (local bits
// bit0..bit0+nbits-1 of r), for first stage index s0.  pos(r) gives the global
// position of local element r.  Returns the accumulated inverse scale.
template <bool INV, int N, int S0, int NST, int BIT0, typename PosFn>
NSB_DEV int fx_local_stages(uint32_t (&v)[8], const uint32_t* tw, bool act, PosFn pos) {
  int scale = 0;
#pragma unroll
  for (int t = 0; t < NST; ++t) {
    const int s = S0 + t;          // stage: span l = 2^s
    int shift = 0;
    if (INV) {
      // complex_fft.c:186-198: scan all 2N values before every stage
      int m = act ? fx_lane_max_abs(v) : 0;
      m = warp_max_i(m);
      if (m > 32767) m = 32767;
      if (m > 13573) ++shift;
      if (m > 27146) ++shift;
      scale += shift;
    }
    if (act) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (!((r >> (BIT0 + t)) & 1)) {
          const int p = pos(r);
          const int tix = (p & ((1 << s) - 1)) << (7 - s);
          fx_butterfly<INV>(v[r], v[r | (1 << (BIT0 + t))], tw[tix], shift);
        }
      }
    }
  }
  return scale;
}

// Full transform. in: layout A (v[r] = element at position 8*lane + r, already
// in bit-reversed order, i.e. position p holds sample bitrev(p)).
// out: layout C (v[r] = element lane + LANES*r). scr: N words of scratch.
template <bool INV, int N>
NSB_DEV int fx_warp_cfft(uint32_t (&v)[8], uint32_t* scr, const uint32_t* tw, int lane) {
  constexpr int LANES = N / 8;
  constexpr int STAGES = N == 256 ? 8 : 7;
  const bool act = lane < LANES;
  int scale = 0;
  scale += fx_local_stages<INV, N, 0, 3, 0>(v, tw, act, [&](int r) { return 8 * lane + r; });
  if (act) {
#pragma unroll
    for (int r = 0; r < 8; ++r) scr[fx_swz(8 * lane + r)] = v[r];
  }
  __syncwarp();
  if (act) {
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = scr[fx_swz((lane & 7) + 8 * r + 64 * (lane >> 3))];
  }
  __syncwarp();
  scale += fx_local_stages<INV, N, 3, 3, 0>(v, tw, act,
                                           [&](int r) { return (lane & 7) + 8 * r + 64 * (lane >> 3); });
  if (act) {
#pragma unroll
    for (int r = 0; r < 8; ++r) scr[fx_swz((lane & 7) + 8 * r + 64 * (lane >> 3))] = v[r];
  }
  __syncwarp();
  if (act) {
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = scr[fx_swz(lane + LANES * r)];
  }
  __syncwarp();
  // remaining stages 6..STAGES-1 act on the top local bits of layout C
  scale += fx_local_stages<INV, N, 6, STAGES - 6, (N == 256 ? 1 : 2)>(
      v, tw, act, [&](int r) { return lane + LANES * r; });
  return scale;
}

}  // namespace nsb200

#endif  // AUDIOSIGNALPROCESS_B200_NS_FIXED_CUH_
