/* TEST INFRASTRUCTURE ONLY: exports the scalar fixed-point helpers of spl_fixed.h
 * so that the known-answer values of the reference's own unit tests
 * (common_audio/signal_processing/signal_processing_unittest.cc:55-56,100-112,
 * 144,148-150,539) can be checked from Python. */
#include "spl_fixed.h"

int32_t oracle_sqrt_floor(int32_t v) { return fx_sqrt_floor(v); }
int oracle_norm_w32(int32_t a) { return fx_norm_w32(a); }
int oracle_norm_u32(uint32_t a) { return fx_norm_u32(a); }
int oracle_norm_w16(int16_t a) { return fx_norm_w16(a); }
int oracle_size_in_bits(uint32_t a) { return fx_size_in_bits(a); }
int32_t oracle_div_w32_w16(int32_t n, int16_t d) { return fx_div_w32_w16(n, d); }
int32_t oracle_div_w32_w16_res16(int32_t n, int16_t d) { return fx_div_w32_w16_res16(n, d); }
uint32_t oracle_div_u32_u16(uint32_t n, uint16_t d) { return fx_div_u32_u16(n, d); }
int32_t oracle_mul16_rsft(int16_t a, int16_t b, int c) { return fx_mul16_rsft(a, b, c); }
int32_t oracle_mul16_rsft_round(int16_t a, int16_t b, int c) { return fx_mul16_rsft_round(a, b, c); }
int32_t oracle_energy(const int16_t* v, int n, int* scale) { return fx_energy(v, n, scale); }
int16_t oracle_max_abs16(const int16_t* v, int n) { return fx_max_abs16(v, n); }
int16_t oracle_sat16(int32_t v) { return fx_sat16(v); }
