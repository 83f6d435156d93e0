/* TEST INFRASTRUCTURE ONLY -- never linked into or called from the product.
 *
 * CPU restatement of the fixed-point noise suppressor WebRtcNsx_* (one stream,
 * scalar C).  Written from the reference's algorithm, not copied from it: the
 * structure, names and table construction are ours; the arithmetic is the
 * reference's, step for step, so that int16 output is bit-identical.  Pinned
 * against the compiled reference (oracle/_ref) in tests/test_oracle_pinning.py
 * and against committed golden vectors in tests/golden/.
 *
 * Reference files (root WebRtc_AMP_Port/webrtc/): NS = modules/
 * audio_processing/ns, SPL = common_audio/signal_processing.
 *   state / init / policy       NS/nsx_core.h:22-110, NS/nsx_core.c:630-813
 *   analysis + start-up model   NS/nsx_core.c:523-551, 1183-1419
 *   int16 FFT pair              SPL/real_fft.c:47-102, complex_bit_reverse.c:49,
 *                               complex_fft.c:29-158 (fwd, mode 1), :160-301 (inv)
 *   spectral flatness / diff    NS/nsx_core.c:1021-1083, 1090-1180
 *   quantile noise tracker      NS/nsx_core.c:303-452
 *   parametric noise            NS/nsx_core.c:585-627, 1615-1711
 *   threshold extraction        NS/nsx_core.c:820-1015
 *   speech/noise probability    NS/nsx_core_c.c:26-261
 *   SNR / noise update / filter NS/nsx_core.c:1723-2031
 *   synthesis                   NS/nsx_core.c:455-520, 1421-1500
 *   high bands                  NS/nsx_core.c:2043-2120
 */
#include "ns_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "spl_fixed.h"

#define NBIN_MAX 129
#define ANA_MAX 256

/* ---- tables, built once by formula (each checked against the reference's
 * literal tables in tests/test_tables.py) ---------------------------------- */
static struct {
  int ready;
  int16_t sin1024[1024];   /* SPL/complex_fft_tables.h kSinTable1024: trunc(32767 sin) */
  int16_t win256[256];     /* NS/nsx_core.c:90  kBlocks160w256x, Q14 */
  int16_t win128[128];     /* NS/nsx_core.c:74  kBlocks80w128x */
  int16_t log_tab[9];      /* :28  kLogTable      round(i 256 ln2) */
  int16_t counter_div[201];/* :32  kCounterDiv    round(32768/(i+1)) */
  int16_t log_frac[256];   /* :49  kLogTableFrac  round(256 log2(1+i/256)) */
  int16_t factor1[257];    /* :135 kFactor1Table */
  int16_t factor2[3][257]; /* :170,193,216 kFactor2Aggressiveness1..3 */
  int16_t sum_log_idx[66]; /* :240 */
  int16_t sum_sq_log_idx[66]; /* :254 */
  int16_t log_idx[129];    /* :268 */
  int16_t det_matrix[66];  /* :290 */
} T;

/* NS/nsx_core_c.c:17 kIndicatorTable: 8192*tanh-like map, tabulated values. */
static const int16_t kIndicator[17] = {0,    2017, 3809, 5227, 6258, 6963, 7424, 7718, 7901,
                                       8014, 8084, 8126, 8152, 8168, 8177, 8183, 8187};

static int rnd(double x) { return (int)floor(x + 0.5); }

static void build_tables(void) {
  int i, k;
  const double pi = 3.14159265358979323846;
  if (T.ready) return;
  for (i = 0; i <= 256; ++i) T.sin1024[i] = (int16_t)(int)(32767.0 * sin(2.0 * pi * i / 1024.0));
  for (i = 257; i <= 512; ++i) T.sin1024[i] = T.sin1024[512 - i];
  for (i = 513; i < 1024; ++i) T.sin1024[i] = (int16_t)-T.sin1024[i - 512];
  for (i = 0; i < 256; ++i) {
    double v = i < 96 ? sin(pi * i / 192.0) : (i <= 160 ? 1.0 : sin(pi * (256 - i) / 192.0));
    T.win256[i] = (int16_t)rnd(16384.0 * v);
  }
  for (i = 0; i < 128; ++i) {
    double v = i < 48 ? sin(pi * i / 96.0) : (i <= 80 ? 1.0 : sin(pi * (128 - i) / 96.0));
    T.win128[i] = (int16_t)rnd(16384.0 * v);
  }
  for (i = 0; i < 9; ++i) T.log_tab[i] = (int16_t)rnd(i * 256.0 * log(2.0));
  T.counter_div[0] = 32767;
  for (i = 1; i < 201; ++i) T.counter_div[i] = (int16_t)rnd(32768.0 / (i + 1));
  for (i = 0; i < 256; ++i) T.log_frac[i] = (int16_t)rnd(256.0 * log2(1.0 + i / 256.0));
  for (i = 0; i < 257; ++i) {
    double g = sqrt(i / 256.0), f = 1.0;
    if (g > 0.5) {
      f = 1.0 + 1.3 * (g - 0.5);
      if (g * f > 1.0) f = 1.0 / g;
    }
    T.factor1[i] = (int16_t)(int)(8192.0 * f);
    for (k = 0; k < 3; ++k) {
      const double bound = k == 0 ? 0.25 : (k == 1 ? 0.125 : 0.09);
      double gg = g, f2 = 1.0;
      if (g <= 0.5) {
        if (gg <= bound) gg = bound;
        f2 = 1.0 - 0.3 * (0.5 - gg);
      }
      T.factor2[k][i] = (int16_t)(int)(8192.0 * f2);
    }
  }
  for (i = 1; i < 129; ++i) T.log_idx[i] = (int16_t)rnd(4096.0 * log2((double)i));
  for (i = 1; i < 66; ++i) {
    double s1 = 0, s2 = 0;
    int j;
    for (j = i; j < 129; ++j) {
      s1 += log2((double)j);
      s2 += log2((double)j) * log2((double)j);
    }
    T.sum_log_idx[i] = (int16_t)rnd(32.0 * s1);
    T.sum_sq_log_idx[i] = (int16_t)rnd(4.0 * s2);
    T.det_matrix[i] = (int16_t)rnd((129 - i) * s2 - s1 * s1);
  }
  T.ready = 1;
}

const int16_t* nsx_oracle_table(const char* name, int* len) {
  build_tables();
#define TAB(n, f) if (!strcmp(name, n)) { *len = (int)(sizeof(T.f) / sizeof(int16_t)); return (const int16_t*)T.f; }
  TAB("kSinTable1024", sin1024) TAB("kBlocks160w256x", win256) TAB("kBlocks80w128x", win128)
  TAB("WebRtcNsx_kLogTable", log_tab) TAB("WebRtcNsx_kCounterDiv", counter_div)
  TAB("WebRtcNsx_kLogTableFrac", log_frac) TAB("kFactor1Table", factor1)
  TAB("kFactor2Aggressiveness1", factor2[0]) TAB("kFactor2Aggressiveness2", factor2[1])
  TAB("kFactor2Aggressiveness3", factor2[2]) TAB("kSumLogIndex", sum_log_idx)
  TAB("kSumSquareLogIndex", sum_sq_log_idx) TAB("kLogIndex", log_idx)
  TAB("kDeterminantEstMatrix", det_matrix)
#undef TAB
  if (!strcmp(name, "kIndicatorTable")) { *len = 17; return kIndicator; }
  *len = 0;
  return NULL;
}

/* ---- state ---------------------------------------------------------------- */
struct NsxOracle {
  int fs, frame, ana, half, nbin, stages, mode, inited;
  const int16_t* window;
  uint16_t overdrive, denoise_bound;
  int gain_map;
  int32_t max_lrt, min_lrt;
  int16_t ana_buf[ANA_MAX], syn_buf[ANA_MAX];
  uint16_t filter[NBIN_MAX];
  int16_t lq[3 * NBIN_MAX], density[3 * NBIN_MAX], quantile[NBIN_MAX], counter[3];
  int32_t lrt_avg[NBIN_MAX];
  int32_t feat_lrt, thr_lrt;
  int16_t w_lrt, w_diff, w_flat;
  uint32_t feat_diff, thr_diff, feat_flat, thr_flat;
  int32_t pause[NBIN_MAX];
  uint32_t magn_energy, sum_magn, cur_avg_energy, time_avg_energy, time_avg_energy_tmp;
  uint32_t white_level, init_magn[NBIN_MAX];
  int32_t pink_num, pink_exp;
  int min_norm, zero_input;
  uint32_t prev_noise[NBIN_MAX];
  uint16_t prev_magn[NBIN_MAX];
  int16_t prior_nonspeech;
  int block_index, cnt_thr_update;
  int16_t hist_lrt[1000], hist_flat[1000], hist_diff[1000];
  int16_t hb_buf[2][ANA_MAX];
  int q_noise, prev_q_noise, prev_q_magn;
  int16_t re[ANA_MAX], im[ANA_MAX];
  int32_t energy_in;
  int scale_energy_in, norm_data;
};

NsxOracle* nsx_oracle_create(void) {
  NsxOracle* s = (NsxOracle*)calloc(1, sizeof(NsxOracle));
  build_tables();
  return s;
}
void nsx_oracle_free(NsxOracle* s) { free(s); }

int nsx_oracle_set_policy(NsxOracle* s, int mode) {
  static const uint16_t od[4] = {256, 256, 282, 320};
  static const uint16_t db[4] = {8192, 4096, 2048, 1475};
  if (!s || mode < 0 || mode > 3) return -1;
  s->mode = mode;
  s->overdrive = od[mode];
  s->denoise_bound = db[mode];
  s->gain_map = mode != 0;
  return 0;
}

int nsx_oracle_init(NsxOracle* s, uint32_t fs) {
  int i;
  if (!s) return -1;
  if (!(fs == 8000 || fs == 16000 || fs == 32000 || fs == 48000)) return -1;
  memset(s, 0, sizeof(*s));
  s->fs = (int)fs;
  if (fs == 8000) {
    s->frame = 80; s->ana = 128; s->stages = 7; s->window = T.win128;
    s->thr_lrt = 131072; s->max_lrt = 0x0040000; s->min_lrt = 52429;
  } else {
    s->frame = 160; s->ana = 256; s->stages = 8; s->window = T.win256;
    s->thr_lrt = 212644; s->max_lrt = 0x0080000; s->min_lrt = 104858;
  }
  s->half = s->ana / 2;
  s->nbin = s->half + 1;
  for (i = 0; i < 3 * NBIN_MAX; ++i) { s->lq[i] = 2048; s->density[i] = 153; }
  for (i = 0; i < 3; ++i) s->counter[i] = (int16_t)((200 * (i + 1)) / 3);
  for (i = 0; i < NBIN_MAX; ++i) s->filter[i] = 16384;
  s->prior_nonspeech = 8192;
  s->thr_diff = 50;
  s->thr_flat = 20480;
  s->feat_lrt = s->thr_lrt;
  s->feat_flat = s->thr_flat;
  s->feat_diff = s->thr_diff;
  s->w_lrt = 6;
  s->block_index = -1;
  s->min_norm = 15;
  nsx_oracle_set_policy(s, 0);
  s->inited = 1;
  return 0;
}

/* ---- int16 complex FFT (n = 2^stages points, interleaved re/im) ------------ */
static void bit_reverse(int16_t* c, int stages) {
  const int n = 1 << stages;
  int i;
  for (i = 0; i < n; ++i) {
    int r = 0, b;
    for (b = 0; b < stages; ++b) r |= ((i >> b) & 1) << (stages - 1 - b);
    if (r > i) {
      int16_t tr = c[2 * i], ti = c[2 * i + 1];
      c[2 * i] = c[2 * r]; c[2 * i + 1] = c[2 * r + 1];
      c[2 * r] = tr; c[2 * r + 1] = ti;
    }
  }
}

/* forward, "mode 1" rounding (complex_fft.c:92-155) */
static void cfft_fwd(int16_t* c, int stages) {
  const int n = 1 << stages;
  int l = 1, k = 9;
  while (l < n) {
    const int step = l << 1;
    int m, i;
    for (m = 0; m < l; ++m) {
      const int j0 = m << k;
      const int16_t wr = T.sin1024[j0 + 256], wi = (int16_t)-T.sin1024[j0];
      for (i = m; i < n; i += step) {
        const int j = i + l;
        int32_t tr = (fx_mul16(wr, c[2 * j]) - fx_mul16(wi, c[2 * j + 1]) + 1) >> 1;
        int32_t ti = (fx_mul16(wr, c[2 * j + 1]) + fx_mul16(wi, c[2 * j]) + 1) >> 1;
        int32_t qr = fx_shl32(c[2 * i], 14), qi = fx_shl32(c[2 * i + 1], 14);
        c[2 * j] = (int16_t)((qr - tr + 16384) >> 15);
        c[2 * j + 1] = (int16_t)((qi - ti + 16384) >> 15);
        c[2 * i] = (int16_t)((qr + tr + 16384) >> 15);
        c[2 * i + 1] = (int16_t)((qi + ti + 16384) >> 15);
      }
    }
    --k;
    l = step;
  }
}

/* inverse with per-stage data-dependent scaling (complex_fft.c:160-301, mode 1);
 * returns the number of right shifts applied */
static int cfft_inv(int16_t* c, int stages) {
  const int n = 1 << stages;
  int l = 1, k = 9, scale = 0;
  while (l < n) {
    const int step = l << 1;
    int shift = 0, m, i;
    int32_t round2 = 8192;
    const int32_t mx = fx_max_abs16(c, 2 * n);
    if (mx > 13573) { shift++; scale++; round2 <<= 1; }
    if (mx > 27146) { shift++; scale++; round2 <<= 1; }
    for (m = 0; m < l; ++m) {
      const int j0 = m << k;
      const int16_t wr = T.sin1024[j0 + 256], wi = T.sin1024[j0];
      for (i = m; i < n; i += step) {
        const int j = i + l;
        int32_t tr = (fx_mul16(wr, c[2 * j]) - fx_mul16(wi, c[2 * j + 1]) + 1) >> 1;
        int32_t ti = (fx_mul16(wr, c[2 * j + 1]) + fx_mul16(wi, c[2 * j]) + 1) >> 1;
        int32_t qr = fx_shl32(c[2 * i], 14), qi = fx_shl32(c[2 * i + 1], 14);
        c[2 * j] = (int16_t)((qr - tr + round2) >> (shift + 14));
        c[2 * j + 1] = (int16_t)((qi - ti + round2) >> (shift + 14));
        c[2 * i] = (int16_t)((qr + tr + round2) >> (shift + 14));
        c[2 * i + 1] = (int16_t)((qi + ti + round2) >> (shift + 14));
      }
    }
    --k;
    l = step;
  }
  return scale;
}

/* real_fft.c:47: n real -> first n+2 int16 of the complex spectrum */
static void real_fft_fwd(int stages, const int16_t* in, int16_t* out) {
  int16_t c[2 * ANA_MAX];
  const int n = 1 << stages;
  int i;
  for (i = 0; i < n; ++i) { c[2 * i] = in[i]; c[2 * i + 1] = 0; }
  bit_reverse(c, stages);
  cfft_fwd(c, stages);
  memcpy(out, c, sizeof(int16_t) * (size_t)(n + 2));
}
/* real_fft.c:74: n+2 int16 of half spectrum -> n real, returns scale */
static int real_fft_inv(int stages, const int16_t* in, int16_t* out) {
  int16_t c[2 * ANA_MAX];
  const int n = 1 << stages;
  int i, sc;
  memcpy(c, in, sizeof(int16_t) * (size_t)(n + 2));
  for (i = n + 2; i < 2 * n; i += 2) {
    c[i] = in[2 * n - i];
    c[i + 1] = (int16_t)-in[2 * n - i + 1];
  }
  bit_reverse(c, stages);
  sc = cfft_inv(c, stages);
  for (i = 0; i < n; ++i) out[i] = c[2 * i];
  return sc;
}

int nsx_oracle_real_fft(int order, int inverse, const int16_t* in, int16_t* out) {
  build_tables();
  if (order < 1 || order > 8) return -1;
  if (inverse) return real_fft_inv(order, in, out);
  real_fft_fwd(order, in, out);
  return 0;
}

/* Q8 log2 of a 16-bit magnitude via the fraction table (nsx_core.c:361-367) */
static int16_t log2_q8(uint32_t v) {
  const int zeros = fx_norm_u32(v);
  const int frac = (int)(((v << zeros) & 0x7FFFFFFFu) >> 23);
  return (int16_t)(((31 - zeros) << 8) + T.log_frac[frac]);
}

/* ---- analysis (nsx_core.c:1183-1419) --------------------------------------- */
static void analyze(NsxOracle* s, const int16_t* frame, uint16_t* magn) {
  int16_t win[ANA_MAX], norm[ANA_MAX], spec[ANA_MAX + 2];
  int i, net_norm, rs_magn, rs_init;
  int16_t max_abs;

  memmove(s->ana_buf, s->ana_buf + s->frame, sizeof(int16_t) * (size_t)(s->ana - s->frame));
  memcpy(s->ana_buf + s->ana - s->frame, frame, sizeof(int16_t) * (size_t)s->frame);
  for (i = 0; i < s->ana; ++i) win[i] = (int16_t)fx_mul16_rsft_round(s->window[i], s->ana_buf[i], 14);

  s->energy_in = fx_energy(win, s->ana, &s->scale_energy_in);
  s->zero_input = 0;
  max_abs = fx_max_abs16(win, s->ana);
  s->norm_data = fx_norm_w16(max_abs);
  if (max_abs == 0) {
    s->zero_input = 1;
    return;
  }
  net_norm = s->stages - s->norm_data;
  rs_magn = s->norm_data - s->min_norm;
  rs_init = -rs_magn > 0 ? -rs_magn : 0;
  s->min_norm -= rs_init;
  if (rs_magn < 0) rs_magn = 0;

  for (i = 0; i < s->ana; ++i) norm[i] = (int16_t)fx_shl32(win[i], s->norm_data);
  real_fft_fwd(s->stages, norm, spec);

  s->im[0] = 0;
  s->im[s->half] = 0;
  s->re[0] = spec[0];
  s->re[s->half] = spec[s->ana];
  s->magn_energy = (uint32_t)(s->re[0] * s->re[0]);
  s->magn_energy += (uint32_t)(s->re[s->half] * s->re[s->half]);
  magn[0] = (uint16_t)(s->re[0] >= 0 ? s->re[0] : -s->re[0]);
  magn[s->half] = (uint16_t)(s->re[s->half] >= 0 ? s->re[s->half] : -s->re[s->half]);
  s->sum_magn = (uint32_t)magn[0] + (uint32_t)magn[s->half];

  if (s->block_index >= 50) {
    for (i = 1; i < s->half; ++i) {
      uint32_t e;
      s->re[i] = spec[2 * i];
      s->im[i] = (int16_t)-spec[2 * i + 1];
      e = (uint32_t)(spec[2 * i] * spec[2 * i]) + (uint32_t)(spec[2 * i + 1] * spec[2 * i + 1]);
      s->magn_energy += e;
      magn[i] = (uint16_t)fx_sqrt_floor((int32_t)e);
      s->sum_magn += magn[i];
    }
    return;
  }
  /* start-up: accumulate the initial spectrum and the pink-noise fit */
  {
    int32_t sum_log_magn, sum_log_i_log_magn, t1, t2;
    int16_t l2 = 0, det, sum_log_i, sum_log_i_sq;
    uint16_t slm_u16, tu16;
    uint32_t tu;
    int zeros;

    s->init_magn[0] >>= rs_init;
    s->init_magn[s->half] >>= rs_init;
    s->init_magn[0] += magn[0] >> rs_magn;
    s->init_magn[s->half] += magn[s->half] >> rs_magn;
    if (magn[s->half]) l2 = log2_q8(magn[s->half]);
    sum_log_magn = l2;
    sum_log_i_log_magn = (T.log_idx[s->half] * l2) >> 3;
    for (i = 1; i < s->half; ++i) {
      uint32_t e;
      s->re[i] = spec[2 * i];
      s->im[i] = (int16_t)-spec[2 * i + 1];
      e = (uint32_t)(spec[2 * i] * spec[2 * i]) + (uint32_t)(spec[2 * i + 1] * spec[2 * i + 1]);
      s->magn_energy += e;
      magn[i] = (uint16_t)fx_sqrt_floor((int32_t)e);
      s->sum_magn += magn[i];
      s->init_magn[i] >>= rs_init;
      s->init_magn[i] += magn[i] >> rs_magn;
      if (i >= 5) {
        l2 = magn[i] ? log2_q8(magn[i]) : 0;
        sum_log_magn += l2;
        sum_log_i_log_magn += (T.log_idx[i] * l2) >> 3;
      }
    }
    s->white_level >>= rs_init;
    tu = fx_umul_32_16(s->sum_magn, s->overdrive);
    tu >>= s->stages + 8;
    tu >>= rs_magn;
    s->white_level += tu;

    det = T.det_matrix[5];
    sum_log_i = T.sum_log_idx[5];
    sum_log_i_sq = T.sum_sq_log_idx[5];
    if (s->fs == 8000) {
      t1 = det;
      t1 += fx_mul16_rsft(T.sum_log_idx[65], sum_log_i, 9);
      t1 -= fx_mul16_rsft(T.sum_log_idx[65], T.sum_log_idx[65], 10);
      t1 -= fx_shl32(sum_log_i_sq, 4);
      t1 -= fx_mul16_rsft((int16_t)(s->nbin - 5), T.sum_sq_log_idx[65], 2);
      det = (int16_t)t1;
      sum_log_i = (int16_t)(sum_log_i - T.sum_log_idx[65]);
      sum_log_i_sq = (int16_t)(sum_log_i_sq - T.sum_sq_log_idx[65]);
    }
    zeros = 16 - fx_norm_w32(sum_log_magn);
    if (zeros < 0) zeros = 0;
    t1 = fx_shl32(sum_log_magn, 1);
    slm_u16 = (uint16_t)(t1 >> zeros);

    t2 = (int32_t)sum_log_i_sq * (int32_t)slm_u16;
    tu = (uint32_t)(sum_log_i_log_magn >> 12);
    tu16 = (uint16_t)((uint16_t)sum_log_i << 1);
    if ((uint32_t)sum_log_i > tu) tu16 >>= zeros;
    else tu >>= zeros;
    t2 -= (int32_t)fx_umul_32_16(tu, tu16);
    det >>= zeros;
    t2 = fx_div_w32_w16(t2, det);
    t2 += fx_shl32(net_norm, 11);
    if (t2 < 0) t2 = 0;
    s->pink_num += t2;

    t2 = (int32_t)sum_log_i * (int32_t)slm_u16;
    t1 = sum_log_i_log_magn >> (3 + zeros);
    t1 *= s->nbin - 5;
    t2 -= t1;
    if (t2 > 0) {
      t1 = fx_div_w32_w16(t2, det);
      s->pink_exp += t1 > 16384 ? 16384 : (t1 < 0 ? 0 : t1);
    }
  }
}

/* ---- synthesis (nsx_core.c:1421-1500, 455-520) ------------------------------ */
static void flush_synthesis(NsxOracle* s, int16_t* out) {
  memcpy(out, s->syn_buf, sizeof(int16_t) * (size_t)s->frame);
  memmove(s->syn_buf, s->syn_buf + s->frame, sizeof(int16_t) * (size_t)(s->ana - s->frame));
  memset(s->syn_buf + s->ana - s->frame, 0, sizeof(int16_t) * (size_t)s->frame);
}

static void synthesize(NsxOracle* s, int16_t* out) {
  int16_t spec[ANA_MAX + 2], time[ANA_MAX];
  int i, sc, scale_out = 0;
  int16_t gain = 8192;
  if (s->zero_input) {
    flush_synthesis(s, out);
    return;
  }
  for (i = 0; i < s->nbin; ++i) {
    s->re[i] = (int16_t)fx_mul16_rsft(s->re[i], (int16_t)s->filter[i], 14);
    s->im[i] = (int16_t)fx_mul16_rsft(s->im[i], (int16_t)s->filter[i], 14);
  }
  for (i = 0; i <= s->half; ++i) {
    spec[2 * i] = s->re[i];
    spec[2 * i + 1] = (int16_t)-s->im[i];
  }
  sc = real_fft_inv(s->stages, spec, time);
  for (i = 0; i < s->ana; ++i) s->re[i] = fx_sat16(fx_shift_w32(time[i], sc - s->norm_data));

  if (s->gain_map == 1 && s->block_index > 200 && s->energy_in > 0) {
    int32_t e_out = fx_energy(s->re, s->ana, &scale_out);
    int16_t ratio, g1, g2;
    int32_t r32;
    if (scale_out == 0 && !(e_out & 0x7f800000)) {
      e_out = fx_shift_w32(e_out, 8 + scale_out - s->scale_energy_in);
    } else {
      s->energy_in >>= 8 + scale_out - s->scale_energy_in;
    }
    r32 = (e_out + s->energy_in / 2) / s->energy_in;
    ratio = (int16_t)r32;                                   /* int16 assignment as in the reference */
    ratio = (int16_t)(ratio > 256 ? 256 : (ratio < 0 ? 0 : ratio));
    g1 = T.factor1[ratio];
    g2 = T.factor2[s->mode - 1][ratio];
    gain = (int16_t)(fx_mul16_rsft((int16_t)(16384 - s->prior_nonspeech), g1, 14) +
                     fx_mul16_rsft(s->prior_nonspeech, g2, 14));
  }
  for (i = 0; i < s->ana; ++i) {
    int16_t a = (int16_t)fx_mul16_rsft_round(s->window[i], s->re[i], 14);
    int32_t t = fx_mul16_rsft_round(a, gain, 13);
    s->syn_buf[i] = fx_sat16((int32_t)s->syn_buf[i] + (int32_t)fx_sat16(t));
  }
  flush_synthesis(s, out);
}

/* ---- features --------------------------------------------------------------- */
static void spectral_flatness(NsxOracle* s, const uint16_t* magn) {
  uint32_t num = 0, den = s->sum_magn - (uint32_t)magn[0];
  int32_t t, lcur, cur;
  int i, zeros, frac, int_part;
  for (i = 1; i < s->nbin; ++i) {
    if (magn[i]) {
      num += (uint32_t)log2_q8(magn[i]);
    } else {
      uint32_t d = fx_umul_32_16(s->feat_flat, 4915);
      s->feat_flat -= d >> 14;
      return;
    }
  }
  zeros = fx_norm_u32(den);
  frac = (int)(((den << zeros) & 0x7FFFFFFFu) >> 23);
  t = (int32_t)(((31 - zeros) << 8) + T.log_frac[frac]);
  lcur = (int32_t)num;
  lcur += fx_shl32(s->stages - 1, s->stages + 7);
  lcur -= fx_shl32(t, s->stages - 1);
  lcur = fx_shl32(lcur, 10 - s->stages);
  t = (int32_t)(0x00020000 | ((lcur >= 0 ? lcur : -lcur) & 0x0001FFFF));
  int_part = 7 - (lcur >> 17);
  cur = int_part > 0 ? (t >> int_part) : fx_shl32(t, -int_part);
  t = cur - (int32_t)s->feat_flat;
  t *= 4915;
  s->feat_flat += (uint32_t)(t >> 14);
}

static void spectral_difference(NsxOracle* s, const uint16_t* magn) {
  int32_t avg_pause = 0, max_p = 0, min_p = s->pause[0], avg_magn, cov = 0, t1, t2;
  uint32_t var_m = 0, var_p = 0, diff, u1, u2;
  int i, n_shifts, norm32;
  for (i = 0; i < s->nbin; ++i) {
    avg_pause += s->pause[i];
    if (s->pause[i] > max_p) max_p = s->pause[i];
    if (s->pause[i] < min_p) min_p = s->pause[i];
  }
  avg_pause >>= s->stages - 1;
  avg_magn = (int32_t)(s->sum_magn >> (s->stages - 1));
  t1 = (max_p - avg_pause) > (avg_pause - min_p) ? (max_p - avg_pause) : (avg_pause - min_p);
  n_shifts = 10 + s->stages - fx_norm_w32(t1);
  if (n_shifts < 0) n_shifts = 0;
  for (i = 0; i < s->nbin; ++i) {
    int16_t d16 = (int16_t)((int32_t)magn[i] - avg_magn);
    t2 = s->pause[i] - avg_pause;
    var_m += (uint32_t)(d16 * d16);
    cov += (int32_t)((uint32_t)t2 * (uint32_t)(int32_t)d16);
    t1 = t2 >> n_shifts;
    var_p += (uint32_t)t1 * (uint32_t)t1;
  }
  s->cur_avg_energy += s->magn_energy >> (2 * s->norm_data + s->stages - 1);
  diff = var_m;
  if (var_p && cov) {
    u1 = (uint32_t)(cov >= 0 ? cov : -cov);
    norm32 = fx_norm_u32(u1) - 16;
    if (norm32 > 0) u1 <<= norm32;
    else u1 >>= -norm32;
    u2 = u1 * u1;
    n_shifts += norm32;
    n_shifts <<= 1;
    if (n_shifts < 0) {
      var_p >>= -n_shifts;
      n_shifts = 0;
    }
    if (var_p > 0) {
      u1 = u2 / var_p;
      u1 >>= n_shifts;
      diff -= diff < u1 ? diff : u1;
    } else {
      diff = 0;
    }
  }
  u1 = diff >> (2 * s->norm_data);
  if (s->feat_diff > u1) {
    u2 = fx_umul_32_16(s->feat_diff - u1, 77);
    s->feat_diff -= u2 >> 8;
  } else {
    u2 = fx_umul_32_16(u1 - s->feat_diff, 77);
    s->feat_diff += u2 >> 8;
  }
}

/* two highest histogram peaks, sequential scan (nsx_core.c:923-939) */
static void two_peaks(const int16_t* h, int* w1, uint32_t* p1, int* w2, uint32_t* p2) {
  int i, m1 = 0, m2 = 0;
  *w1 = *w2 = 0;
  *p1 = *p2 = 0;
  for (i = 0; i < 1000; ++i) {
    if (h[i] > m1) {
      m2 = m1; *w2 = *w1; *p2 = *p1;
      m1 = h[i]; *w1 = h[i]; *p1 = (uint32_t)(2 * i + 1);
    } else if (h[i] > m2) {
      m2 = h[i]; *w2 = h[i]; *p2 = (uint32_t)(2 * i + 1);
    }
  }
}

static void feature_extraction(NsxOracle* s, int flag) {
  int i;
  if (!flag) {
    uint32_t idx = (uint32_t)s->feat_lrt;
    if (idx < 1000) s->hist_lrt[idx]++;
    idx = (s->feat_flat * 5) >> 8;
    if (idx < 1000) s->hist_flat[idx]++;
    idx = 1000;
    if (s->time_avg_energy > 0) idx = ((s->feat_diff * 5) >> s->stages) / s->time_avg_energy;
    if (idx < 1000) s->hist_diff[idx]++;
    return;
  }
  {
    int use_diff = 1, use_flat, w1, w2, fsum;
    uint32_t p1, p2, tu;
    int32_t avg = 0, avg_sq = 0, avg_compl, fluct, thr_fluct, t;
    int16_t num = 0;
    for (i = 0; i < 10; ++i) {
      int16_t j = (int16_t)(2 * i + 1);
      t = s->hist_lrt[i] * j;
      avg += t;
      num = (int16_t)(num + s->hist_lrt[i]);
      avg_sq += t * j;
    }
    avg_compl = avg;
    for (; i < 1000; ++i) {
      int16_t j = (int16_t)(2 * i + 1);
      t = s->hist_lrt[i] * j;
      avg_compl += t;
      avg_sq += t * j;
    }
    fluct = avg_sq * num - avg * avg_compl;
    thr_fluct = 10240 * num;
    tu = 6u * (uint32_t)avg;
    if (fluct < thr_fluct || num == 0 || tu > (uint32_t)(100 * num)) {
      s->thr_lrt = s->max_lrt;
    } else {
      t = (int32_t)((tu << (9 + s->stages)) / (uint32_t)num / 25);
      s->thr_lrt = t > s->max_lrt ? s->max_lrt : (t < s->min_lrt ? s->min_lrt : t);
    }
    if (fluct < thr_fluct) use_diff = 0;

    two_peaks(s->hist_flat, &w1, &p1, &w2, &p2);
    use_flat = 1;
    if ((p1 - p2 < 4) && (w2 * 2 > w1)) {
      w1 += w2;
      p1 = (p1 + p2) >> 1;
    }
    if (w1 < 154 || p1 < 24) {
      use_flat = 0;
    } else {
      uint32_t v = 922u * p1;
      s->thr_flat = v > 38912 ? 38912 : (v < 4096 ? 4096 : v);
    }
    if (use_diff) {
      two_peaks(s->hist_diff, &w1, &p1, &w2, &p2);
      if ((p1 - p2 < 4) && (w2 * 2 > w1)) {
        w1 += w2;
        p1 = (p1 + p2) >> 1;
      }
      {
        uint32_t v = 6u * p1;
        s->thr_diff = v > 100 ? 100 : (v < 16 ? 16 : v);
      }
      if (w1 < 154) use_diff = 0;
    }
    fsum = 6 / (1 + use_flat + use_diff);
    s->w_lrt = (int16_t)fsum;
    s->w_flat = (int16_t)(use_flat * fsum);
    s->w_diff = (int16_t)(use_diff * fsum);
    memset(s->hist_lrt, 0, sizeof(s->hist_lrt));
    memset(s->hist_flat, 0, sizeof(s->hist_flat));
    memset(s->hist_diff, 0, sizeof(s->hist_diff));
  }
}

/* ---- quantile noise tracker (nsx_core.c:303-452) ----------------------------- */
static void latch_quantile(NsxOracle* s, int offset) {
  int i;
  int16_t mx = fx_max16(s->lq + offset, s->nbin);
  s->q_noise = 14 - (int)fx_mul16_rsft_round(11819, mx, 21);
  for (i = 0; i < s->nbin; ++i) {
    int32_t e = 11819 * s->lq[offset + i];
    int32_t m = 0x00200000 | (e & 0x001FFFFF);
    int16_t sh = (int16_t)(e >> 21);
    sh = (int16_t)(sh - 21);
    sh = (int16_t)(sh + s->q_noise);
    if (sh < 0) m >>= -sh;
    else m = fx_shl32(m, sh);
    s->quantile[i] = fx_sat16(m);
  }
}

static void noise_estimation(NsxOracle* s, const uint16_t* magn, uint32_t* noise, int16_t* q_noise) {
  int16_t lmagn[NBIN_MAX], logval;
  int i, k, offset = 0;
  int tabind = s->stages - s->norm_data;
  logval = tabind < 0 ? (int16_t)-T.log_tab[-tabind] : T.log_tab[tabind];
  for (i = 0; i < s->nbin; ++i) {
    if (magn[i]) {
      int16_t l2 = log2_q8(magn[i]);
      lmagn[i] = (int16_t)fx_mul16_rsft(l2, 22713, 15);
      lmagn[i] = (int16_t)(lmagn[i] + logval);
    } else {
      lmagn[i] = logval;
    }
  }
  for (k = 0; k < 3; ++k) {
    const int16_t counter = s->counter[k];
    const int16_t cdiv = T.counter_div[counter];
    const int16_t cprod = (int16_t)(counter * cdiv);
    offset = k * s->nbin;
    for (i = 0; i < s->nbin; ++i) {
      int16_t delta, t16;
      int16_t* lq = &s->lq[offset + i];
      int16_t* dn = &s->density[offset + i];
      if (*dn > 512) {
        delta = (int16_t)(2621440 >> (14 - fx_norm_w16(*dn)));
      } else {
        delta = s->block_index < 200 ? 1024 : 5120;
      }
      t16 = (int16_t)fx_mul16_rsft(delta, cdiv, 14);
      if (lmagn[i] > *lq) {
        t16 = (int16_t)(t16 + 2);
        *lq = (int16_t)(*lq + t16 / 4);
      } else {
        int16_t t2;
        t16 = (int16_t)(t16 + 1);
        t2 = (int16_t)fx_mul16_rsft((int16_t)(t16 / 2), 3, 1);
        *lq = (int16_t)(*lq - t2);
        if (*lq < logval) *lq = logval;
      }
      {
        int16_t d = (int16_t)(lmagn[i] - *lq);
        if ((d >= 0 ? d : -d) < 3) {
          int16_t a = (int16_t)fx_mul16_rsft_round(*dn, cprod, 15);
          int16_t b = (int16_t)fx_mul16_rsft_round(21845, cdiv, 15);
          *dn = (int16_t)(a + b);
        }
      }
    }
    if (counter >= 200) {
      s->counter[k] = 0;
      if (s->block_index >= 200) latch_quantile(s, offset);
    }
    s->counter[k]++;
  }
  if (s->block_index < 200) latch_quantile(s, offset);
  for (i = 0; i < s->nbin; ++i) noise[i] = (uint32_t)s->quantile[i];
  *q_noise = (int16_t)s->q_noise;
}

/* 2^(num - exp*log2(bin)) in Q(minNorm-stages) (nsx_core.c:585-627).  Leaves the
 * outputs untouched when the exponent is not positive, like the reference. */
static void parametric_noise(const NsxOracle* s, int16_t exp_avg, int32_t num_avg, int bin,
                             uint32_t* est, uint32_t* est_avg) {
  int32_t t2 = (exp_avg * T.log_idx[bin]) >> 15;
  int32_t t1 = num_avg - t2;
  t1 += fx_shl32(s->min_norm - s->stages, 11);
  if (t1 > 0) {
    const int int_part = (int16_t)(t1 >> 11);
    const int frac = (int16_t)(t1 & 0x7ff);
    if (frac >> 10) {
      t2 = (2048 - frac) * 1244;
      t2 = 2048 - (t2 >> 10);
    } else {
      t2 = (frac * 804) >> 10;
    }
    t2 = fx_shift_w32(t2, int_part - 11);
    *est_avg = (uint32_t)fx_shl32(1, int_part) + (uint32_t)t2;
    *est = *est_avg * (uint32_t)(s->block_index + 1);
  }
}

/* ---- speech / noise probability (nsx_core_c.c:26-261) ------------------------ */
static int16_t indicator_interp(uint32_t x_q14, int rounded) {
  const int idx = (int16_t)(x_q14 >> 14);
  const int16_t base = kIndicator[idx];
  const int16_t step = (int16_t)(kIndicator[idx + 1] - kIndicator[idx]);
  const int16_t frac = (int16_t)(x_q14 & 0x3fff);
  return (int16_t)(base + (int16_t)(rounded ? fx_mul16_rsft_round(step, frac, 14)
                                            : fx_mul16_rsft(step, frac, 14)));
}

static void speech_noise_prob(NsxOracle* s, uint16_t* nonspeech, const uint32_t* prior_snr,
                              const uint32_t* post_snr) {
  int32_t ksum = 0, ind_prior, t1;
  int16_t ind, ind16, d16;
  int i, n_shifts;
  for (i = 0; i < s->nbin; ++i) {
    int32_t bessel = (int32_t)post_snr[i], frac32, t, log_t, half_sum;
    const int nt = fx_norm_u32(post_snr[i]);
    const uint32_t num = post_snr[i] << nt;
    const uint32_t den = nt > 10 ? (prior_snr[i] << (nt - 11)) : (prior_snr[i] >> (11 - nt));
    int zeros;
    if (den > 0) bessel -= (int32_t)(num / den);
    else bessel = 0;
    zeros = fx_norm_u32(prior_snr[i]);
    frac32 = (int32_t)(((prior_snr[i] << zeros) & 0x7FFFFFFFu) >> 19);
    t = (frac32 * frac32 * -43) >> 19;
    t += fx_mul16_rsft((int16_t)frac32, 5412, 12);
    frac32 = t + 37;
    t = (int32_t)(((31 - zeros) << 12) + frac32) - (11 << 12);
    log_t = (t * 178) >> 8;
    half_sum = (log_t + s->lrt_avg[i]) / 2;
    s->lrt_avg[i] += bessel - half_sum;
    ksum += s->lrt_avg[i];
  }
  s->feat_lrt = (ksum * 10) >> (s->stages + 11);

  /* LRT indicator */
  ind = 16384;
  t1 = ksum - s->thr_lrt;
  n_shifts = 7 - s->stages;
  if (t1 < 0) {
    ind = 0;
    t1 = -t1;
    n_shifts++;
  }
  t1 = fx_shift_w32(t1, n_shifts);
  {
    const int16_t idx = (int16_t)(t1 >> 14);
    if (idx < 16 && idx >= 0) {
      const int16_t v = indicator_interp((uint32_t)t1, 0);
      ind = (int16_t)(ind == 0 ? 8192 - v : 8192 + v);
    }
  }
  ind_prior = s->w_lrt * ind;

  if (s->w_flat) {
    uint32_t u1 = s->feat_flat * 400u, u2;
    ind = 16384;
    u2 = s->thr_flat - u1;
    n_shifts = 4;
    if (s->thr_flat < u1) {
      ind = 0;
      u2 = u1 - s->thr_flat;
      n_shifts++;
    }
    u1 = fx_div_u32_u16(u2 << n_shifts, 25);
    if ((int16_t)(u1 >> 14) < 16) {
      const int16_t v = indicator_interp(u1, 0);
      ind = (int16_t)(ind ? 8192 + v : 8192 - v);
    }
    ind_prior += s->w_flat * ind;
  }
  if (s->w_diff) {
    uint32_t u1 = 0, u2, u3;
    if (s->feat_diff) {
      int nt = fx_norm_u32(s->feat_diff);
      if (20 - s->stages < nt) nt = 20 - s->stages;
      u1 = s->feat_diff << nt;
      u2 = s->time_avg_energy >> (20 - s->stages - nt);
      if (u2 > 0) u1 /= u2;
      else u1 = 0x7fffffffu;
    }
    u3 = (s->thr_diff << 17) / 25;
    u2 = u1 - u3;
    n_shifts = 1;
    ind = 16384;
    if (u2 & 0x80000000u) {
      ind = 0;
      u2 = u3 - u1;
      n_shifts--;
    }
    u1 = u2 >> n_shifts;
    if ((int16_t)(u1 >> 14) < 16) {
      const int16_t v = indicator_interp(u1, 1);
      ind = (int16_t)(ind ? 8192 + v : 8192 - v);
    }
    ind_prior += s->w_diff * ind;
  }
  ind16 = fx_div_w32_w16_res16(98307 - ind_prior, 6);
  d16 = (int16_t)(ind16 - s->prior_nonspeech);
  s->prior_nonspeech = (int16_t)(s->prior_nonspeech + (int16_t)fx_mul16_rsft(1638, d16, 14));

  memset(nonspeech, 0, sizeof(uint16_t) * (size_t)s->nbin);
  if (s->prior_nonspeech > 0) {
    for (i = 0; i < s->nbin; ++i) {
      if (s->lrt_avg[i] < 65300) {
        int32_t e = (s->lrt_avg[i] * 23637) >> 14, t2, inv;
        int16_t int_part = (int16_t)(e >> 12), frac;
        int n1, n2;
        if (int_part < -8) int_part = -8;
        frac = (int16_t)(e & 0xfff);
        t2 = (frac * frac * 44) >> 19;
        t2 += fx_mul16_rsft(frac, 84, 7);
        inv = fx_shl32(1, 8 + int_part) + fx_shift_w32(t2, int_part - 4);
        n1 = fx_norm_w32(inv);
        n2 = fx_norm_w16((int16_t)(16384 - s->prior_nonspeech));
        if (n1 + n2 >= 7) {
          int32_t p;
          if (n1 + n2 < 15) {
            inv >>= 15 - n2 - n1;
            p = inv * (16384 - s->prior_nonspeech);
            inv = fx_shift_w32(p, 7 - n1 - n2);
          } else {
            p = inv * (16384 - s->prior_nonspeech);
            inv = p >> 8;
          }
          p = fx_shl32(s->prior_nonspeech, 8);
          nonspeech[i] = (uint16_t)(p / (s->prior_nonspeech + inv));
        }
      }
    }
  }
}

/* ---- one frame (nsx_core.c:1502-2121) ---------------------------------------- */
static void hb_push(NsxOracle* s, int b, const int16_t* in) {
  memmove(s->hb_buf[b], s->hb_buf[b] + s->frame, sizeof(int16_t) * (size_t)(s->ana - s->frame));
  memcpy(s->hb_buf[b] + s->ana - s->frame, in, sizeof(int16_t) * (size_t)s->frame);
}

void nsx_oracle_process(NsxOracle* s, const int16_t* const* in, int num_bands, int16_t* const* out) {
  uint16_t magn[NBIN_MAX], prev_noise16[NBIN_MAX], nonspeech[NBIN_MAX], filter_tmp[NBIN_MAX];
  uint32_t noise[NBIN_MAX], post_snr[NBIN_MAX], prior_snr[NBIN_MAX], prev_near[NBIN_MAX];
  uint32_t max_noise, u1, u2, u3;
  const uint32_t sat_max = 1048575u;
  int16_t q_magn, q_noise;
  int i, b, n_shifts, post_shifts, norm1, flag;
  uint16_t gamma, prev_gamma;

  analyze(s, in[0], magn);
  if (s->zero_input) {
    synthesize(s, out[0]);
    for (b = 0; b < num_bands - 1; ++b) {
      hb_push(s, b, in[b + 1]);
      memcpy(out[b + 1], s->hb_buf[b], sizeof(int16_t) * (size_t)s->frame);
    }
    return;
  }
  s->block_index++;
  q_magn = (int16_t)(s->norm_data - s->stages);
  spectral_flatness(s, magn);
  noise_estimation(s, magn, noise, &q_noise);
  for (i = 0; i < s->nbin; ++i) prev_noise16[i] = (uint16_t)(s->prev_noise[i] >> 11);

  if (s->block_index < 50) {
    uint32_t est = 0, est_avg = 0, numer;
    int32_t num_avg = 0;
    int16_t exp_avg = 0;
    int q_use = (int)q_noise < s->min_norm - s->stages ? (int)q_noise : s->min_norm - s->stages;
    if (s->pink_exp) {
      exp_avg = (int16_t)fx_div_w32_w16(s->pink_exp, (int16_t)(s->block_index + 1));
      num_avg = fx_div_w32_w16(s->pink_num, (int16_t)(s->block_index + 1));
      parametric_noise(s, exp_avg, num_avg, 5, &est, &est_avg);
    } else {
      est = s->white_level;
      est_avg = est / (uint32_t)(s->block_index + 1);
    }
    for (i = 0; i < s->nbin; ++i) {
      if (s->pink_exp && i >= 5) {
        est = 0;
        est_avg = 0;
        parametric_noise(s, exp_avg, num_avg, i, &est, &est_avg);
      }
      filter_tmp[i] = s->denoise_bound;
      if (s->init_magn[i]) {
        u1 = fx_umul_32_16(est, s->overdrive);
        numer = s->init_magn[i] << 8;
        if (numer > u1) {
          int ns;
          numer -= u1;
          ns = fx_norm_u32(numer);
          ns = ns > 6 ? 6 : (ns < 0 ? 0 : ns);
          numer <<= ns;
          u1 = s->init_magn[i] >> (6 - ns);
          if (u1 == 0) u1 = 1;
          u2 = numer / u1;
          filter_tmp[i] = (uint16_t)(u2 > 16384 ? 16384 : (u2 < (uint32_t)s->denoise_bound ? s->denoise_bound : u2));
        }
      }
      u1 = noise[i] >> (q_noise - q_use);
      u2 = est_avg >> (s->min_norm - s->stages - q_use);
      n_shifts = 0;
      if (u1 & 0xfc000000u) {
        u1 >>= 6;
        u2 >>= 6;
        n_shifts = 6;
      }
      u1 *= (uint32_t)s->block_index;
      u2 *= (uint32_t)(50 - s->block_index);
      noise[i] = fx_div_u32_u16(u1 + u2, 50);
      noise[i] <<= n_shifts;
    }
    q_noise = (int16_t)q_use;
  }
  if (s->block_index < 200) {
    s->time_avg_energy_tmp += s->magn_energy >> (2 * s->norm_data + s->stages - 1);
    s->time_avg_energy = fx_div_u32_u16(s->time_avg_energy_tmp, (uint16_t)(s->block_index + 1));
  }

  /* step 1: post / prior SNR */
  post_shifts = 6 + q_magn - q_noise;
  n_shifts = 5 - s->prev_q_magn + s->prev_q_noise;
  for (i = 0; i < s->nbin; ++i) {
    uint32_t near_est, prior;
    post_snr[i] = 2048;
    u1 = (uint32_t)magn[i] << 6;
    u2 = post_shifts < 0 ? (noise[i] >> -post_shifts) : (noise[i] << post_shifts);
    if (u1 > u2) {
      u1 <<= 11;
      if (u2 > 0) {
        u1 /= u2;
        post_snr[i] = u1 < sat_max ? u1 : sat_max;
      } else {
        post_snr[i] = sat_max;
      }
    }
    near_est = (uint32_t)s->prev_magn[i] * (uint32_t)s->filter[i];
    u1 = near_est << 3;
    u2 = s->prev_noise[i] >> n_shifts;
    if (u2 > 0) {
      u1 /= u2;
      if (u1 > sat_max) u1 = sat_max;
    } else {
      u1 = sat_max;
    }
    prev_near[i] = u1;
    u1 = fx_umul_32_16(prev_near[i], 2007);
    u2 = fx_umul_32_16(post_snr[i] - 2048, 41);
    prior = u1 + u2 + 512;
    prior_snr[i] = 2048 + (prior >> 10);
  }

  /* step 2: features, probability, noise update */
  spectral_difference(s, magn);
  s->cnt_thr_update++;
  flag = s->cnt_thr_update == 512;
  feature_extraction(s, flag);
  if (flag) {
    s->cnt_thr_update = 0;
    s->cur_avg_energy >>= 9;
    u1 = (s->cur_avg_energy + s->time_avg_energy + 1) >> 1;
    if (u1 != s->time_avg_energy && s->feat_diff && s->time_avg_energy > 0) {
      norm1 = 0;
      u3 = u1;
      while (0xFFFF0000u & u3) { u3 >>= 1; norm1++; }
      u2 = s->feat_diff;
      while (0xFFFF0000u & u2) { u2 >>= 1; norm1++; }
      u3 = u3 * u2;
      u3 /= s->time_avg_energy;
      if (fx_norm_u32(u3) < norm1) {
        s->feat_diff = 0x007FFFFF;
      } else {
        uint32_t v = u3 << norm1;
        s->feat_diff = v < 0x007FFFFF ? v : 0x007FFFFF;
      }
    }
    s->time_avg_energy = u1;
    s->cur_avg_energy = 0;
  }
  speech_noise_prob(s, nonspeech, prior_snr, post_snr);

  gamma = 26;
  max_noise = 0;
  post_shifts = s->prev_q_noise - q_magn;
  n_shifts = s->prev_q_magn - q_magn;
  for (i = 0; i < s->nbin; ++i) {
    uint32_t upd;
    int sign;
    int32_t t1, t2;
    u2 = post_shifts < 0 ? ((uint32_t)magn[i] >> -post_shifts) : ((uint32_t)magn[i] << post_shifts);
    if (prev_noise16[i] > u2) { sign = -1; u1 = prev_noise16[i] - u2; }
    else { sign = 1; u1 = u2 - prev_noise16[i]; }
    upd = s->prev_noise[i];
    u3 = 0;
    if (u1 && nonspeech[i]) {
      u3 = fx_umul_32_16(u1, nonspeech[i]);
      u2 = (0x7c000000u & u3) ? (u3 >> 5) * gamma : (u3 * gamma) >> 5;
      if (sign > 0) upd += u2;
      else upd -= u2;
    }
    prev_gamma = gamma;
    gamma = nonspeech[i] < 205 ? 3 : 26;
    if (prev_gamma != gamma) {
      u2 = (0x7c000000u & u3) ? (u3 >> 5) * gamma : (u3 * gamma) >> 5;
      u1 = sign > 0 ? s->prev_noise[i] + u2 : s->prev_noise[i] - u2;
      if (upd > u1) upd = u1;
    }
    noise[i] = upd;
    if (upd > max_noise) max_noise = upd;
    t2 = fx_shift_w32(s->pause[i], -n_shifts);
    if (nonspeech[i] > 205) {
      if (n_shifts < 0) {
        t1 = (int32_t)magn[i] - t2;
        t1 *= 13;
        t1 = (t1 + 128) >> 8;
      } else {
        t1 = fx_shl32((int32_t)magn[i], n_shifts) - s->pause[i];
        t1 *= 13;
        t1 = (t1 + fx_shl32(128, n_shifts)) >> (8 + n_shifts);
      }
      t2 += t1;
    }
    s->pause[i] = t2;
  }
  norm1 = fx_norm_u32(max_noise);
  q_noise = (int16_t)(s->prev_q_noise + norm1 - 5);

  /* step 3: Wiener filter from the updated noise */
  n_shifts = s->prev_q_noise + 11 - q_magn;
  for (i = 0; i < s->nbin; ++i) {
    uint32_t cur = 0, tm, tn, prior;
    uint16_t f16;
    if (n_shifts < 0) {
      tm = magn[i];
      tn = noise[i] << -n_shifts;
    } else if (n_shifts > 17) {
      tm = (uint32_t)magn[i] << 17;
      tn = noise[i] >> (n_shifts - 17);
    } else {
      tm = (uint32_t)magn[i] << n_shifts;
      tn = noise[i];
    }
    if (tm > tn) {
      int nn;
      u1 = tm - tn;
      nn = fx_norm_u32(u1);
      if (nn > 11) nn = 11;
      u1 <<= nn;
      u2 = tn >> (11 - nn);
      if (u2 > 0) u1 /= u2;
      cur = u1 < sat_max ? u1 : sat_max;
    }
    u1 = fx_umul_32_16(prev_near[i], 2007);
    u2 = fx_umul_32_16(cur, 41);
    prior = u1 + u2;
    u1 = (uint32_t)s->overdrive + ((prior + 8192) >> 14);
    f16 = (uint16_t)((prior + u1 / 2) / u1);
    s->filter[i] = (uint16_t)(f16 > 16384 ? 16384 : (f16 < s->denoise_bound ? s->denoise_bound : f16));
    if (s->block_index < 50) {
      u1 = (uint32_t)s->filter[i] * (uint32_t)s->block_index;
      u2 = (uint32_t)filter_tmp[i] * (uint32_t)(50 - s->block_index);
      s->filter[i] = (uint16_t)fx_div_u32_u16(u1 + u2, 50);
    }
  }
  s->prev_q_noise = q_noise;
  s->prev_q_magn = q_magn;
  for (i = 0; i < s->nbin; ++i) {
    s->prev_noise[i] = norm1 > 5 ? noise[i] << (norm1 - 5) : noise[i] >> (5 - norm1);
    s->prev_magn[i] = magn[i];
  }
  synthesize(s, out[0]);

  if (num_bands > 1) {
    int16_t g = 16384, avg_prob, gain_mod, avg_gain;
    uint16_t psum = 0;
    for (b = 0; b < num_bands - 1; ++b) hb_push(s, b, in[b + 1]);
    u1 = 0;
    for (i = s->half - (s->half >> 2); i < s->half; ++i) {
      psum = (uint16_t)(psum + nonspeech[i]);
      u1 += s->filter[i];
    }
    avg_prob = (int16_t)(4096 - (psum >> (s->stages - 7)));
    avg_gain = (int16_t)(u1 >> (s->stages - 3));
    gain_mod = avg_prob < 3607 ? avg_prob : 3607;
    if (avg_prob < 2048) {
      g = (int16_t)((gain_mod << 1) + (avg_gain >> 1));
    } else {
      g = (int16_t)fx_mul16_rsft(3, avg_gain, 2);
      g = (int16_t)(g + gain_mod);
    }
    g = (int16_t)(g > 16384 ? 16384 : (g < (int16_t)s->denoise_bound ? (int16_t)s->denoise_bound : g));
    for (b = 0; b < num_bands - 1; ++b)
      for (i = 0; i < s->frame; ++i) out[b + 1][i] = (int16_t)fx_mul16_rsft(g, s->hb_buf[b][i], 14);
  }
}

int nsx_oracle_run(int fs, int mode, int nframes, const int16_t* pcm_in, int16_t* pcm_out) {
  NsxOracle* s = nsx_oracle_create();
  int f;
  if (!s || nsx_oracle_init(s, (uint32_t)fs) != 0 || nsx_oracle_set_policy(s, mode) != 0) {
    nsx_oracle_free(s);
    return -1;
  }
  if (fs > 16000) {  /* band split lives in band_oracle.c; single-band entry only here */
    nsx_oracle_free(s);
    return -2;
  }
  for (f = 0; f < nframes; ++f) {
    const int16_t* in[1] = {pcm_in + (size_t)f * s->frame};
    int16_t* out[1] = {pcm_out + (size_t)f * s->frame};
    nsx_oracle_process(s, in, 1, out);
  }
  nsx_oracle_free(s);
  return 0;
}
