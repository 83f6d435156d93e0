/* TEST INFRASTRUCTURE ONLY (oracle): scalar fixed-point helpers restating the
 * WebRTC signal-processing-library primitives the NSx path uses.  Each function
 * cites the reference definition it follows; arithmetic is bit-identical, the
 * code is our own.  Reference root: WebRtc_AMP_Port/webrtc/common_audio/
 * signal_processing/ (abbrev. SPL/).
 */
#ifndef ORACLE_SPL_FIXED_H_
#define ORACLE_SPL_FIXED_H_

#include <stdint.h>
#include <stdlib.h>

/* SPL/include/spl_inl.h:27 WebRtcSpl_SatW32ToW16 */
static inline int16_t fx_sat16(int32_t v) {
  return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
}
/* spl_inl.h:103 WebRtcSpl_NormW32: redundant sign bits of a (0 for a == 0) */
static inline int fx_norm_w32(int32_t a) {
  if (a == 0) return 0;
  if (a < 0) a = ~a;
  return a == 0 ? 31 : __builtin_clz((uint32_t)a) - 1;
}
/* spl_inl.h:126 WebRtcSpl_NormU32: leading zeros (0 for a == 0) */
static inline int fx_norm_u32(uint32_t a) { return a == 0 ? 0 : __builtin_clz(a); }
/* spl_inl.h:144 WebRtcSpl_NormW16 */
static inline int fx_norm_w16(int16_t a) {
  int32_t v = a;
  if (v == 0) return 0;
  if (v < 0) v = ~v;
  return v == 0 ? 15 : __builtin_clz((uint32_t)v) - 17;
}
/* spl_inl.h:86 WebRtcSpl_GetSizeInBits */
static inline int fx_size_in_bits(uint32_t n) { return n == 0 ? 0 : 32 - __builtin_clz(n); }

/* SPL/include/signal_processing_library.h:49-101 multiply/shift macros */
static inline int32_t fx_mul16(int16_t a, int16_t b) { return (int32_t)a * (int32_t)b; }
static inline int32_t fx_mul16_rsft(int16_t a, int16_t b, int c) { return fx_mul16(a, b) >> c; }
static inline int32_t fx_mul16_rsft_round(int16_t a, int16_t b, int c) {
  return (fx_mul16(a, b) + ((int32_t)1 << (c - 1))) >> c;
}
static inline uint32_t fx_umul_32_16(uint32_t a, uint16_t b) { return a * (uint32_t)b; }
/* WEBRTC_SPL_SHIFT_W32: left for c >= 0 else arithmetic right.  Left shifts are
 * done on the unsigned image (two's-complement wrap, what the reference's
 * compilers produce). */
static inline int32_t fx_shift_w32(int32_t x, int c) {
  return c >= 0 ? (int32_t)((uint32_t)x << c) : (x >> (-c));
}
static inline int32_t fx_shl32(int32_t x, int c) { return (int32_t)((uint32_t)x << c); }

/* SPL/division_operations.c:26,38,50: zero-guarded divides */
static inline uint32_t fx_div_u32_u16(uint32_t num, uint16_t den) {
  return den ? num / den : 0xFFFFFFFFu;
}
static inline int32_t fx_div_w32_w16(int32_t num, int16_t den) {
  return den ? num / den : 0x7FFFFFFF;
}
static inline int16_t fx_div_w32_w16_res16(int32_t num, int16_t den) {
  return den ? (int16_t)(num / den) : (int16_t)0x7FFF;
}

/* SPL/spl_sqrt_floor.c:55: floor(sqrt(value)) by 16 restoring iterations;
 * negative input yields 0 (no iteration can fire). */
static inline int32_t fx_sqrt_floor(int32_t value) {
  int32_t root = 0;
  int n;
  for (n = 15; n >= 0; --n) {
    int32_t t = root + ((int32_t)1 << n);
    if (value >= (int32_t)((uint32_t)t << n)) {
      value -= (int32_t)((uint32_t)t << n);
      root |= (int32_t)2 << n;
    }
  }
  return root >> 1;
}

/* SPL/min_max_operations.c:36 WebRtcSpl_MaxAbsValueW16C (|-32768| clamps to 32767) */
static inline int16_t fx_max_abs16(const int16_t* v, int n) {
  int m = 0, i;
  for (i = 0; i < n; ++i) {
    int a = abs((int)v[i]);
    if (a > m) m = a;
  }
  return (int16_t)(m > 32767 ? 32767 : m);
}
/* min_max_operations.c:84 WebRtcSpl_MaxValueW16C */
static inline int16_t fx_max16(const int16_t* v, int n) {
  int16_t m = -32768;
  int i;
  for (i = 0; i < n; ++i)
    if (v[i] > m) m = v[i];
  return m;
}

/* SPL/get_scaling_square.c:20 + SPL/energy.c:20: energy with the right shift
 * that keeps the sum inside int32. */
static inline int32_t fx_energy(const int16_t* v, int n, int* scale) {
  int nbits = fx_size_in_bits((uint32_t)n);
  int smax = -1, i, sh;
  int32_t en = 0;
  for (i = 0; i < n; ++i) {
    int16_t s = v[i] > 0 ? v[i] : (int16_t)-v[i];  /* int16 negate: -(-32768) stays -32768 */
    if (s > smax) smax = s;
  }
  if (smax == 0) {
    sh = 0;
  } else {
    int t = fx_norm_w32((int32_t)smax * (int32_t)smax);
    sh = t > nbits ? 0 : nbits - t;
  }
  for (i = 0; i < n; ++i) en += fx_mul16(v[i], v[i]) >> sh;
  *scale = sh;
  return en;
}

#endif /* ORACLE_SPL_FIXED_H_ */
