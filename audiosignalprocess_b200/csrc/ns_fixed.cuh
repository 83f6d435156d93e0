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

// floor(a / b) for b >= 1 when the quotient is below 1 195 000 (1.14 * 2^20); a value of at least 2^20 - 1
// otherwise.  Every division in the per-bin loops either has a quotient bounded far below that (bounds at the
// call sites) or is capped at 2^20 - 1 right after, as the reference caps it.  The estimate is a float product
// built to never exceed the true quotient -- divisor rounded up, reciprocal (1 ulp) scaled by 1 - 2^-22,
// dividend and product rounded toward zero: between (a/b)(1 - 7 * 2^-23) and a/b -- so its floor is the
// quotient or one less while a/b < 2^20.19, the remainder a - q b cannot wrap, and one comparison finishes.
// 9 instructions, four of them conversions / MUFU off the integer pipe, against ~19 for the compiler's general
// 32-bit division.  Checked on the device against `/` by WebRtcNsB200_SelfTest (random operands and
// operands placed one below, at and one above exact multiples).
constexpr unsigned kFxUdivExactBelow = 1195000u;
NSB_DEV unsigned fx_udiv_q20(unsigned a, unsigned b) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__uint2float_ru(b)));
  r = __fmul_rz(r, 0.99999976158142089844f);   // 1 - 2^-22
  const unsigned q = __float2uint_rz(__fmul_rz(__uint2float_rz(a), r));
  return a - q * b >= b ? q + 1u : q;
#else
  return a / b;
#endif
}

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
// N = 256 (8 stages) or 128 (7 stages) complex points, LANES = N/8 active lanes, 8 points per lane.
// Radix-2 decimation in time in exactly the reference's stage order (complex_fft.c:29-158 forward, :160-301
// inverse, mode 1) with its truncating int16 store after every stage; three register layouts, two
// conflict-free shared-memory transposes:
//   A: position p = 8*lane + r                      stages 0,1,2
//   B: p = (lane&7) + 8*r + 64*(lane>>3)            stages 3,4,5
//   C: p = lane + LANES*r                           stages 6(,7)
// The points stay UNPACKED in registers between the stages of a layout (re[8], im[8] as sign-extended
// ints): the reference's `(int16_t)(x >> 15)` is bits 15..30 of x, i.e. the upper half of 2x, so with every
// term of the butterfly doubled the truncating store is one arithmetic shift by 16, and nothing is packed or
// unpacked around a stage (16 instructions per forward butterfly where the packed form took ~40).  Words
// are packed (re low, im high) only to cross a transpose.
//
// Twiddles (cos, sin)(2 pi t / 256) = (kSinTable1024[4t + 256], kSinTable1024[4t]) arrive as int2 in a table
// regrouped per stage so that the lanes of a warp read consecutive entries (the butterflies of stage s use
// t = (p mod 2^s) << (7 - s): strided, up to 8-way bank conflicts straight from a linear table):
//   [0, 4)            stages 0-2: t = 32 q
//   [4 + 8 g, ..+8)   stages 3-5, g = 0 (s=3) | 1 + (r&1) (s=4) | 3 + (r&3) (s=5): t for lane&7 = 0..7
//   [60, 124)         stage 6: N=256: (r&1)*32 + lane;  N=128: (r&3)*16 + lane
//   [124, 252)        stage 7 (N=256): (r&3)*32 + lane = t itself
constexpr int kFxTwInt2 = 252;   // filled on the host: nsx_host_init.h nsx_fill_fft_twiddles

NSB_DEV int fx_swz(int p) { return p ^ ((p >> 3) & 7) ^ (((p >> 6) & 3) << 3); }

// (re, im) int16 pairs in one word.  On the device each is ONE byte permute (the compiler's own sequences are
// two instructions for the pack and for the sign-extending low half), on the integer ALU pipe this kernel is
// bound by; selector 0x9910 = bytes 0, 1 and the sign of byte 1 replicated.
NSB_DEV uint32_t fx_pack(int re, int im) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(r) : "r"(re), "r"(im));
  return r;
#else
  return ((uint32_t)re & 0xffffu) | ((uint32_t)im << 16);
#endif
}
NSB_DEV int fx_lo(uint32_t w) {
#ifdef __CUDA_ARCH__
  int r;
  asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(r) : "r"(w));
  return r;
#else
  return (int)(int16_t)(w & 0xffffu);
#endif
}
NSB_DEV int fx_hi(uint32_t w) { return (int)(int16_t)(w >> 16); }

// One butterfly on unpacked points a (kept) and x (twiddled), w = (cos, sin).
// Forward: t = ((wr xr + ws xi + 1) >> 1, (wr xi - ws xr + 1) >> 1); out = (int16)((a 2^14 -+ t + 2^14) >> 15).
// Inverse: t = ((wr xr - ws xi + 1) >> 1, (wr xi + ws xr + 1) >> 1); out = (int16)((a 2^14 -+ t + (2^13 << shift)) >> (14 + shift)).
//
// With p = the twiddle product incl. its + 1, floor((A + floor(p / 2)) / 2^k) = floor((2A + p) / 2^(k+1)) and
// floor((A - floor(p / 2)) / 2^k) = floor((2A + 1 - p) / 2^(k+1)) for even A (a dropped or added odd unit never
// crosses a multiple of an even divisor), so the halving disappears and each output is ONE multiply-add chain
//   a * 2^15 + rounding (+ 1) +- wr xr +- ws xi
// followed by one shift: the integer ALU pipe (shifts, adds, logic: half the rate of the multiply-add pipe, and
// the pipe this kernel saturates) sees 4 instructions per butterfly instead of 12.  The int16 cast is the
// upper half of the sum scaled to put the kept bits at 16..31 (wrap-around included: the bits shifted out are
// the ones the cast discards).  The inverse transform at shift 2 keeps bit 31, cannot be doubled, and takes
// the plain form.
// SH = -1: forward transform; SH = 0, 1, 2: inverse transform at that scaling shift.  The shift is decided per
// stage from the data (fx_inverse_shift) and is the same for the whole warp, so the stage is compiled once per
// value and selected by one warp-uniform branch per stage, instead of a branch and run-time scale factors in
// every butterfly.
template <int SH>
NSB_DEV void fx_butterfly_u(int& ar, int& ai, int& xr, int& xi, int2 w) {
  if (SH < 0) {
    const unsigned qr = (unsigned)(ar * 32768 + 32768), qi = (unsigned)(ai * 32768 + 32768);
    const int mr = w.x * xr + w.y * xi, mi = w.x * xi - w.y * xr;     // p - 1
    xr = (int)(qr - (unsigned)mr) >> 16;
    xi = (int)(qi - (unsigned)mi) >> 16;
    ar = (int)(qr + 1u + (unsigned)mr) >> 16;
    ai = (int)(qi + 1u + (unsigned)mi) >> 16;
  } else if (SH < 2) {
    // scaled by m = 2^(1 - shift): a * 2^(16 - shift) + 2^15 (+ m) +- m (wr xr - ws xi)   (wrap-around ring
    // arithmetic: the factor m moves freely between the operands)
    constexpr int m = 2 - SH;                                     // 2 or 1
    const unsigned qr = (unsigned)(ar * (32768 * m) + 32768), qi = (unsigned)(ai * (32768 * m) + 32768);
    const int br = w.x * xr - w.y * xi, bi = w.x * xi + w.y * xr;
    xr = (int)(qr - (unsigned)(br * m)) >> 16;
    xi = (int)(qi - (unsigned)(bi * m)) >> 16;
    ar = (int)(qr + (unsigned)m + (unsigned)(br * m)) >> 16;
    ai = (int)(qi + (unsigned)m + (unsigned)(bi * m)) >> 16;
  } else {
    const int tr = (w.x * xr - w.y * xi + 1) >> 1, ti = (w.x * xi + w.y * xr + 1) >> 1;
    const int qr = ar * 16384 + 32768, qi = ai * 16384 + 32768;
    xr = (qr - tr) >> 16;
    xi = (qi - ti) >> 16;
    ar = (qr + tr) >> 16;
    ai = (qi + ti) >> 16;
  }
}

// The inverse transform's scaling decision before a stage (complex_fft.c:186-198): the largest |value| of all
// 2N int16 (32768 counts as 32767) against 13573 and 27146.
NSB_DEV int fx_inverse_shift(const int (&re)[8], const int (&im)[8], bool act) {
  int mx = 0, mn = 0;
  if (act) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      mx = re[r] > mx ? re[r] : mx;
      mx = im[r] > mx ? im[r] : mx;
      mn = re[r] < mn ? re[r] : mn;
      mn = im[r] < mn ? im[r] : mn;
    }
  }
  int m = -mn > mx ? -mn : mx;
  m = warp_max_i(m);
  return (m > 13573 ? 1 : 0) + (m > 27146 ? 1 : 0);   // (32768 vs 32767: both above either threshold)
}

// Runs NST stages starting at stage S0 on the lane's 8 points; the butterfly partners differ in local bit
// BIT0 + t.  twi(s, r) = index of the butterfly's twiddle in the regrouped table.  Returns the accumulated
// inverse scale.
template <int SH, int STAGE, int BITPOS, typename TwFn>
NSB_DEV void fx_stage_u(int (&re)[8], int (&im)[8], const int2* tw, TwFn twi) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if (!((r >> BITPOS) & 1)) {
      const int r2 = r | (1 << BITPOS);
      fx_butterfly_u<SH>(re[r], im[r], re[r2], im[r2], tw[twi(STAGE, r)]);
    }
  }
}
template <bool INV, int S0, int NST, int BIT0, int T, typename TwFn>
NSB_DEV int fx_local_stages_from(int (&re)[8], int (&im)[8], const int2* tw, bool act, TwFn twi) {
  int scale = 0;
  if (INV) {
    const int shift = fx_inverse_shift(re, im, act);
    scale = shift;
    if (act) {
      if (shift == 0) fx_stage_u<0, S0 + T, BIT0 + T>(re, im, tw, twi);
      else if (shift == 1) fx_stage_u<1, S0 + T, BIT0 + T>(re, im, tw, twi);
      else fx_stage_u<2, S0 + T, BIT0 + T>(re, im, tw, twi);
    }
  } else if (act) {
    fx_stage_u<-1, S0 + T, BIT0 + T>(re, im, tw, twi);
  }
  if (T + 1 < NST) scale += fx_local_stages_from<INV, S0, NST, BIT0, (T + 1 < NST ? T + 1 : T)>(re, im, tw, act, twi);
  return scale;
}
template <bool INV, int S0, int NST, int BIT0, typename TwFn>
NSB_DEV int fx_local_stages_u(int (&re)[8], int (&im)[8], const int2* tw, bool act, TwFn twi) {
  return fx_local_stages_from<INV, S0, NST, BIT0, 0>(re, im, tw, act, twi);
}

// Full transform. in: layout A (re/im[r] = point at position 8*lane + r, already in bit-reversed order, i.e.
// position p holds sample bitrev(p)).  out: layout C (point lane + LANES*r).  scr: 2 N words of scratch, 8-byte
// aligned: the two transposes move (re, im) as one 64-bit word per point -- no packing to int16 pairs and no
// sign-extending unpack on the integer ALU pipe (24 instructions per transpose and lane), the same number of
// shared-memory instructions; the XOR swizzle keeps every half-warp of 64-bit accesses on distinct bank pairs
// (checked per pattern: the low four bits of the swizzled index differ within lanes 0-15 and 16-31).
template <bool INV, int N>
NSB_DEV int fx_warp_cfft(int (&re)[8], int (&im)[8], uint32_t* scr, const int2* tw, int lane) {
  constexpr int LANES = N / 8;
  constexpr int STAGES = N == 256 ? 8 : 7;
  const bool act = lane < LANES;
  const int l7 = lane & 7;
  int2* s2 = reinterpret_cast<int2*>(scr);
  int scale = 0;
  // stages 0-2: t = (r mod 2^s) << (7 - s) = 32 * {0 | 2 (r&1) | r&3}
  scale += fx_local_stages_u<INV, 0, 3, 0>(re, im, tw, act, [&](int s, int r) { return s == 0 ? 0 : (s == 1 ? 2 * (r & 1) : (r & 3)); });
  if (act) {
#pragma unroll
    for (int r = 0; r < 8; ++r) s2[fx_swz(8 * lane + r)] = make_int2(re[r], im[r]);
  }
  __syncwarp();
  if (act) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int2 w = s2[fx_swz(l7 + 8 * r + 64 * (lane >> 3))];
      re[r] = w.x;
      im[r] = w.y;
    }
  }
  __syncwarp();
  // stages 3-5: the butterfly's position mod 2^s is l7 + 8 (r mod 2^(s-3))
  scale += fx_local_stages_u<INV, 3, 3, 0>(re, im, tw, act, [&](int s, int r) {
    return 4 + 8 * (s == 3 ? 0 : (s == 4 ? 1 + (r & 1) : 3 + (r & 3))) + l7;
  });
  if (act) {
#pragma unroll
    for (int r = 0; r < 8; ++r) s2[fx_swz(l7 + 8 * r + 64 * (lane >> 3))] = make_int2(re[r], im[r]);
  }
  __syncwarp();
  if (act) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int2 w = s2[fx_swz(lane + LANES * r)];
      re[r] = w.x;
      im[r] = w.y;
    }
  }
  __syncwarp();
  // remaining stages 6..STAGES-1 act on the top local bits of layout C
  scale += fx_local_stages_u<INV, 6, STAGES - 6, (N == 256 ? 1 : 2)>(re, im, tw, act, [&](int s, int r) {
    return s == 6 ? 60 + (N == 256 ? (r & 1) * 32 : (r & 3) * 16) + lane : 124 + (r & 3) * 32 + lane;
  });
  return scale;
}

}  // namespace nsb200

#endif  // AUDIOSIGNALPROCESS_B200_NS_FIXED_CUH_
