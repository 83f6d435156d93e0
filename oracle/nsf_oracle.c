/* TEST INFRASTRUCTURE ONLY -- never linked into or called from the product.
 *
 * CPU restatement of the float noise suppressor WebRtcNs_Analyze +
 * WebRtcNs_Process (one stream, scalar C), written from the reference's
 * algorithm: same operations in the same order in single precision with the
 * reference's double-precision libm calls, including the rounding order of its
 * real FFT (restated below, not transcribed).  The float suppressor branches on
 * comparisons that are decided by the last bit of log|X[k]|, so "same order"
 * is what parity means here: this oracle is pinned BIT FOR BIT against the
 * compiled reference (oracle/_ref) in tests/test_oracle_pinning.py and against
 * tests/golden/.
 *
 * Reference (root WebRtc_AMP_Port/webrtc/modules/audio_processing/ns/):
 *   state, init, policy            ns_core.h:52-114, ns_core.c:23-214, 1013-1041
 *   WebRtcNs_AnalyzeCore           ns_core.c:1043-1181
 *   WebRtcNs_ProcessCore           ns_core.c:1183-1415
 *   NoiseEstimation :217, FeatureParameterExtraction :293, ComputeSpectralFlatness
 *   :523, ComputeSnr :566, ComputeSpectralDifference :595, SpeechNoiseProb :642,
 *   FeatureUpdate :755, UpdateNoiseEstimate :800, UpdateBuffer :855, FFT :886,
 *   IFFT :923, Energy :951, Windowing :969, ComputeDdBasedWienerFilter :985
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ns_oracle.h"

#define NB_MAX 129
#define ANA_MAX 256

struct NsfOracle {
  int fs, frame, ana, nbin, mode, inited;
  float window[ANA_MAX];
  float analyze_buf[ANA_MAX], data_buf[ANA_MAX], synt_buf[ANA_MAX], hb_buf[2][ANA_MAX];
  float density[3 * NB_MAX], lquantile[3 * NB_MAX], quantile[NB_MAX];
  int counter[3], updates;
  float smooth[NB_MAX], overdrive, denoise_bound;
  int gainmap;
  int block_ind, upd_flag, upd_count;
  float prior_pars[7];
  float noise[NB_MAX], noise_prev[NB_MAX], magn_prev_analyze[NB_MAX], magn_prev_process[NB_MAX];
  float log_lrt[NB_MAX], prior_speech_prob, feat[7], magn_avg_pause[NB_MAX];
  float signal_energy, sum_magn, white_level, init_magn[NB_MAX], pink_num, pink_exp;
  float parametric[NB_MAX], speech_prob[NB_MAX];
  int hist_lrt[1000], hist_flat[1000], hist_diff[1000];
  float fw_re[ANA_MAX / 8], fw_im[ANA_MAX / 8];   /* quarter-circle twiddles, bit-reversed order */
  float fc[ANA_MAX / 4 + 1];                      /* half-cosine table of the real split */
};

static void fft_tables(NsfOracle* s);

NsfOracle* nsf_oracle_create(void) { return (NsfOracle*)calloc(1, sizeof(NsfOracle)); }
void nsf_oracle_free(NsfOracle* s) { free(s); }

int nsf_oracle_set_policy(NsfOracle* s, int mode) {
  static const float od[4] = {1.f, 1.f, 1.1f, 1.25f};
  static const float db[4] = {0.5f, 0.25f, 0.125f, 0.09f};
  if (!s || mode < 0 || mode > 3) return -1;
  s->mode = mode;
  s->overdrive = od[mode];
  s->denoise_bound = db[mode];
  s->gainmap = mode != 0;
  return 0;
}

int nsf_oracle_init(NsfOracle* s, uint32_t fs) {
  int i, rise;
  const double pi = 3.14159265358979323846;
  if (!s) return -1;
  if (!(fs == 8000 || fs == 16000 || fs == 32000 || fs == 48000)) return -1;
  memset(s, 0, sizeof(*s));
  s->fs = (int)fs;
  s->frame = fs == 8000 ? 80 : 160;
  s->ana = fs == 8000 ? 128 : 256;
  s->nbin = s->ana / 2 + 1;
  rise = s->ana - s->frame;
  /* hybrid Hann/flat window tabulated to 8 decimals (windows_private.h:64,94) */
  for (i = 0; i < s->ana; ++i) {
    double v = i < rise ? sin(pi * i / (2.0 * rise)) : (i <= s->ana - rise ? 1.0 : sin(pi * (s->ana - i) / (2.0 * rise)));
    s->window[i] = (float)(floor(v * 1e8 + 0.5) / 1e8);
  }
  fft_tables(s);
  for (i = 0; i < 3 * NB_MAX; ++i) {
    s->lquantile[i] = 8.f;
    s->density[i] = 0.3f;
  }
  for (i = 0; i < 3; ++i) s->counter[i] = (int)floor((float)(200 * (i + 1)) / 3.f);
  for (i = 0; i < NB_MAX; ++i) {
    s->smooth[i] = 1.f;
    s->log_lrt[i] = 0.5f;
  }
  s->prior_speech_prob = 0.5f;
  s->feat[0] = 0.5f;
  s->feat[3] = 0.5f;
  s->feat[4] = 0.5f;
  s->block_ind = -1;
  s->prior_pars[0] = 0.5f;
  s->prior_pars[1] = 0.5f;
  s->prior_pars[2] = 1.f;
  s->prior_pars[3] = 0.5f;
  s->prior_pars[4] = 1.f;
  s->upd_flag = 2;
  s->upd_count = 500;
  nsf_oracle_set_policy(s, 0);
  s->inited = 1;
  return 0;
}

float nsf_oracle_prior_speech_probability(const NsfOracle* s) {
  return (!s || !s->inited) ? -1.f : s->prior_speech_prob;
}

/* ---- real FFT pair in the operation order of the reference's Ooura rdft ---------------------
 * (utility/fft4g.c:324-361 as ns_core.c:886-944 calls it).  Float addition and multiplication are
 * not associative, and the noise tracker branches on log|X[k]| to the last bit, so this oracle
 * reproduces the reference's transform rounding for rounding; it is restated, not transcribed:
 *   - the n/2 complex points are put in bit-reversed order (what bitrv2, fft4g.c:693, amounts to);
 *   - radix-4 passes over strides 1, 4, 16 (cft1st :1002 is the stride-1 case of cftmdl :1107): the
 *     four points of a butterfly are x[j + {0,1,2,3} * stride]; group g = j / (4 * stride) decides
 *     the three twiddles applied to the butterfly's OUTPUTS (tw_group below), with the sums formed
 *     as (a0 + a1) + (a2 + a3) etc. exactly as the reference forms them;
 *   - a last pass without twiddles: radix-2 for 128 complex points, radix-4 for 64 (cftfsub :902,
 *     cftbsub :952 -- the inverse runs the same twiddled passes on the conjugate and conjugates
 *     back inside this last pass);
 *   - the real-input split (rftfsub :1234 / rftbsub :1259) with the half-cosine table of makect :671.
 * Twiddles: W[q], q < n/8, is the quarter-circle table of makewt (:642) -- (cos, sin)(pi q / (n/4)) with
 * the angle formed in float as the reference forms it -- read in bit-reversed order.
 * Pinned bit for bit against WebRtc_rdft itself (oracle/_ref hook ref_rdft) in tests/test_oracle_pinning.py. */
static unsigned bit_reverse(unsigned v, int bits) {
  unsigned r = 0;
  int b;
  for (b = 0; b < bits; ++b) r |= ((v >> b) & 1u) << (bits - 1 - b);
  return r;
}
static int ilog2(int v) {
  int b = 0;
  while ((1 << b) < v) ++b;
  return b;
}

static void fft_tables(NsfOracle* s) {
  const int n = s->ana, nw = n >> 2, nwh = nw >> 1, nq = nw >> 1, qbits = ilog2(nq);
  const float delta = (float)atan(1.0) / (float)nwh;
  float w[ANA_MAX / 4];
  int j, q;
  w[0] = 1.f;
  w[1] = 0.f;
  w[nwh] = (float)cos((double)(delta * (float)nwh));
  w[nwh + 1] = w[nwh];
  for (j = 2; j < nwh; j += 2) {
    const float ang = delta * (float)j;
    const float x = (float)cos((double)ang), y = (float)sin((double)ang);
    w[j] = x;
    w[j + 1] = y;
    w[nw - j] = y;
    w[nw - j + 1] = x;
  }
  for (q = 0; q < nq; ++q) {
    const unsigned r = bit_reverse((unsigned)q, qbits);
    s->fw_re[q] = w[2 * r];
    s->fw_im[q] = w[2 * r + 1];
  }
  /* half-cosine table of the real split: c[j] = cos(pi j / (n/2)) / 2, c[nc - j] = sin(...) / 2 */
  {
    const int nc = n >> 2, nch = nc >> 1;
    const float d2 = (float)atan(1.0) / (float)nch;
    s->fc[0] = (float)cos((double)(d2 * (float)nch));
    s->fc[nch] = 0.5f * s->fc[0];
    for (j = 1; j < nch; ++j) {
      const float ang = d2 * (float)j;
      s->fc[j] = 0.5f * (float)cos((double)ang);
      s->fc[nc - j] = 0.5f * (float)sin((double)ang);
    }
  }
}

typedef struct { float r1, i1, r2, i2, r3, i3; int diag; } TwGroup;

/* Twiddles of butterfly group g (any radix-4 pass): point 1 (offset stride) gets w1, point 2 w2, point 3 w3.
 * g = 1 is the reference's pi/4 case, which multiplies sums instead of summing products (diag). */
static TwGroup tw_group(const NsfOracle* s, int g) {
  TwGroup t = {1.f, 0.f, 1.f, 0.f, 1.f, 0.f, 0};
  const int p = g >> 1;
  float ar, ai;
  if (g == 0) return t;
  ar = s->fw_re[p];
  ai = s->fw_im[p];
  if ((g & 1) == 0) {
    t.r1 = s->fw_re[2 * p];
    t.i1 = s->fw_im[2 * p];
    t.r2 = ar;
    t.i2 = ai;
    t.r3 = t.r1 - 2.f * ai * t.i1;
    t.i3 = 2.f * ai * t.r1 - t.i1;
  } else {
    t.r1 = s->fw_re[2 * p + 1];
    t.i1 = s->fw_im[2 * p + 1];
    t.r2 = -ai;
    t.i2 = ar;
    t.r3 = t.r1 - 2.f * ar * t.i1;
    t.i3 = 2.f * ar * t.r1 - t.i1;
    t.diag = g == 1;
  }
  return t;
}

/* the twiddled radix-4 passes shared by both directions, in place on bit-reversed data */
static int fft_twiddled_passes(const NsfOracle* s, float* re, float* im, int nc) {
  int l = 1, first = 1;
  while (first || (l << 2) < nc) {
    const int m = l << 2;
    int g, j;
    first = 0;
    for (g = 0; g < nc / m; ++g) {
      const TwGroup t = tw_group(s, g);
      for (j = g * m; j < g * m + l; ++j) {
        const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
        const float x0r = re[j] + re[j1], x0i = im[j] + im[j1];
        const float x1r = re[j] - re[j1], x1i = im[j] - im[j1];
        const float x2r = re[j2] + re[j3], x2i = im[j2] + im[j3];
        const float x3r = re[j2] - re[j3], x3i = im[j2] - im[j3];
        const float dr = x0r - x2r, di = x0i - x2i;
        const float yr = x1r - x3i, yi = x1i + x3r;
        const float zr = x1r + x3i, zi = x1i - x3r;
        re[j] = x0r + x2r;
        im[j] = x0i + x2i;
        re[j2] = t.r2 * dr - t.i2 * di;
        im[j2] = t.r2 * di + t.i2 * dr;
        if (t.diag) {
          re[j1] = t.r1 * (yr - yi);
          im[j1] = t.r1 * (yr + yi);
          re[j3] = t.r1 * (-zi - zr);
          im[j3] = t.r1 * (-zi + zr);
        } else {
          re[j1] = t.r1 * yr - t.i1 * yi;
          im[j1] = t.r1 * yi + t.i1 * yr;
          re[j3] = t.r3 * zr - t.i3 * zi;
          im[j3] = t.r3 * zi + t.i3 * zr;
        }
      }
    }
    l = m;
  }
  return l;
}

/* a[0..n): in = time samples, out = the reference's packed spectrum (a[0] = X[0], a[1] = X[n/2],
 * a[2k], a[2k+1] = Re, Im X[k]; X[k] = sum_j x[j] e^{+2 pi i jk/n}) */
static void rdft_forward(const NsfOracle* s, float* a) {
  const int n = s->ana, nc = n / 2, bits = ilog2(nc), ncq = n >> 2;
  float re[ANA_MAX / 2], im[ANA_MAX / 2];
  int j, l;
  for (j = 0; j < nc; ++j) {
    const unsigned r = bit_reverse((unsigned)j, bits);
    re[j] = a[2 * r];
    im[j] = a[2 * r + 1];
  }
  l = fft_twiddled_passes(s, re, im, nc);
  if ((l << 2) == nc) {
    for (j = 0; j < l; ++j) {
      const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
      const float x0r = re[j] + re[j1], x0i = im[j] + im[j1];
      const float x1r = re[j] - re[j1], x1i = im[j] - im[j1];
      const float x2r = re[j2] + re[j3], x2i = im[j2] + im[j3];
      const float x3r = re[j2] - re[j3], x3i = im[j2] - im[j3];
      re[j] = x0r + x2r;  im[j] = x0i + x2i;
      re[j2] = x0r - x2r; im[j2] = x0i - x2i;
      re[j1] = x1r - x3i; im[j1] = x1i + x3r;
      re[j3] = x1r + x3i; im[j3] = x1i - x3r;
    }
  } else {
    for (j = 0; j < l; ++j) {
      const int j1 = j + l;
      const float dr = re[j] - re[j1], di = im[j] - im[j1];
      re[j] += re[j1];
      im[j] += im[j1];
      re[j1] = dr;
      im[j1] = di;
    }
  }
  for (j = 1; j < nc / 2; ++j) {
    const int k = nc - j;
    const float wkr = 0.5f - s->fc[ncq - j], wki = s->fc[j];
    const float xr = re[j] - re[k], xi = im[j] + im[k];
    const float yr = wkr * xr - wki * xi, yi = wkr * xi + wki * xr;
    re[j] -= yr;
    im[j] -= yi;
    re[k] += yr;
    im[k] -= yi;
  }
  {
    const float d = re[0] - im[0];
    re[0] += im[0];
    im[0] = d;
  }
  for (j = 0; j < nc; ++j) {
    a[2 * j] = re[j];
    a[2 * j + 1] = im[j];
  }
}

/* inverse of the above, unscaled (the caller applies 2/n, ns_core.c:941-943) */
static void rdft_backward(const NsfOracle* s, float* a) {
  const int n = s->ana, nc = n / 2, bits = ilog2(nc), ncq = n >> 2;
  float re[ANA_MAX / 2], im[ANA_MAX / 2], tr[ANA_MAX / 2] = {0.f}, ti[ANA_MAX / 2] = {0.f};
  int j, l;
  for (j = 0; j < nc; ++j) {
    tr[j] = a[2 * j];
    ti[j] = a[2 * j + 1];
  }
  ti[0] = 0.5f * (tr[0] - ti[0]);
  tr[0] -= ti[0];
  ti[0] = -ti[0];
  for (j = 1; j < nc / 2; ++j) {
    const int k = nc - j;
    const float wkr = 0.5f - s->fc[ncq - j], wki = s->fc[j];
    const float xr = tr[j] - tr[k], xi = ti[j] + ti[k];
    const float yr = wkr * xr + wki * xi, yi = wkr * xi - wki * xr;
    tr[j] -= yr;
    ti[j] = yi - ti[j];
    tr[k] += yr;
    ti[k] = yi - ti[k];
  }
  ti[nc / 2] = -ti[nc / 2];
  for (j = 0; j < nc; ++j) {
    const unsigned r = bit_reverse((unsigned)j, bits);
    re[j] = tr[r];
    im[j] = ti[r];
  }
  l = fft_twiddled_passes(s, re, im, nc);
  if ((l << 2) == nc) {
    for (j = 0; j < l; ++j) {
      const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
      const float x0r = re[j] + re[j1], x0i = -im[j] - im[j1];
      const float x1r = re[j] - re[j1], x1i = -im[j] + im[j1];
      const float x2r = re[j2] + re[j3], x2i = im[j2] + im[j3];
      const float x3r = re[j2] - re[j3], x3i = im[j2] - im[j3];
      re[j] = x0r + x2r;  im[j] = x0i - x2i;
      re[j2] = x0r - x2r; im[j2] = x0i + x2i;
      re[j1] = x1r - x3i; im[j1] = x1i - x3r;
      re[j3] = x1r + x3i; im[j3] = x1i + x3r;
    }
  } else {
    for (j = 0; j < l; ++j) {
      const int j1 = j + l;
      const float dr = re[j] - re[j1], di = -im[j] + im[j1];
      re[j] += re[j1];
      im[j] = -im[j] - im[j1];
      re[j1] = dr;
      im[j1] = di;
    }
  }
  for (j = 0; j < nc; ++j) {
    a[2 * j] = re[j];
    a[2 * j + 1] = im[j];
  }
}

/* test hook: the transform alone (n = 256 or 128; isgn >= 0 forward, < 0 backward, unscaled), pinned
 * bit for bit against WebRtc_rdft in tests/test_oracle_pinning.py */
int nsf_oracle_rdft(int n, int isgn, float* a) {
  NsfOracle* s;
  if (n != 256 && n != 128) return -1;
  s = nsf_oracle_create();
  if (!s) return -1;
  s->ana = n;
  fft_tables(s);
  if (isgn >= 0) rdft_forward(s, a);
  else rdft_backward(s, a);
  nsf_oracle_free(s);
  return 0;
}

/* FFT() of ns_core.c:886-911 */
static void rfft_fwd(const NsfOracle* s, float* x, float* xr, float* xi) {
  const int h = s->ana / 2;
  int k;
  rdft_forward(s, x);
  xi[0] = 0.f;
  xr[0] = x[0];
  xi[h] = 0.f;
  xr[h] = x[1];
  for (k = 1; k < h; ++k) {
    xr[k] = x[2 * k];
    xi[k] = x[2 * k + 1];
  }
}

/* IFFT() of ns_core.c:923-944 including the 2/n scaling */
static void rfft_inv(const NsfOracle* s, const float* xr, const float* xi, float* x) {
  const int n = s->ana, h = n / 2;
  int k;
  x[0] = xr[0];
  x[1] = xr[h];
  for (k = 1; k < h; ++k) {
    x[2 * k] = xr[k];
    x[2 * k + 1] = xi[k];
  }
  rdft_backward(s, x);
  for (k = 0; k < n; ++k) x[k] *= 2.f / (float)n;
}

static void push(float* buf, const float* frame, int frame_len, int buf_len) {
  memmove(buf, buf + frame_len, sizeof(float) * (size_t)(buf_len - frame_len));
  if (frame) memcpy(buf + buf_len - frame_len, frame, sizeof(float) * (size_t)frame_len);
  else memset(buf + buf_len - frame_len, 0, sizeof(float) * (size_t)frame_len);
}

static float windowed_energy(const NsfOracle* s, const float* buf, float* win) {
  float e = 0.f;
  int i;
  for (i = 0; i < s->ana; ++i) win[i] = s->window[i] * buf[i];
  for (i = 0; i < s->ana; ++i) e += win[i] * win[i];
  return e;
}

static void spectrum(const NsfOracle* s, float* win, float* re, float* im, float* magn) {
  int i;
  rfft_fwd(s, win, re, im);
  magn[0] = (float)(fabs(re[0]) + 1.f);
  magn[s->nbin - 1] = (float)(fabs(re[s->nbin - 1]) + 1.f);
  for (i = 1; i < s->nbin - 1; ++i) magn[i] = sqrtf(re[i] * re[i] + im[i] * im[i]) + 1.f;
}

static void two_peaks(const int* h, float bin, int* w1, float* p1, int* w2, float* p2) {
  int i, m1 = 0, m2 = 0;
  *w1 = *w2 = 0;
  *p1 = *p2 = 0.f;
  for (i = 0; i < 1000; ++i) {
    const float mid = ((float)i + 0.5f) * bin;
    if (h[i] > m1) {
      m2 = m1; *w2 = *w1; *p2 = *p1;
      m1 = h[i]; *w1 = h[i]; *p1 = mid;
    } else if (h[i] > m2) {
      m2 = h[i]; *w2 = h[i]; *p2 = mid;
    }
  }
}

static void extract_thresholds(NsfOracle* s) {
  float avg = 0.f, avg_c = 0.f, avg_sq = 0.f, fluct, p1, p2, fsum;
  int num = 0, i, w1, w2, use_flat = 1, use_diff = 1;
  for (i = 0; i < 1000; ++i) {
    const float mid = ((float)i + 0.5f) * 0.1f;
    if (mid <= 1.f) {
      avg += s->hist_lrt[i] * mid;
      num += s->hist_lrt[i];
    }
    avg_sq += s->hist_lrt[i] * mid * mid;
    avg_c += s->hist_lrt[i] * mid;
  }
  if (num > 0) avg = avg / (float)num;
  avg_c = avg_c / 500.f;
  avg_sq = avg_sq / 500.f;
  fluct = avg_sq - avg * avg_c;
  if (fluct < 0.05f) {
    s->prior_pars[0] = 1.f;
  } else {
    float t = 1.2f * avg;
    if (t < 0.2f) t = 0.2f;
    if (t > 1.f) t = 1.f;
    s->prior_pars[0] = t;
  }
  two_peaks(s->hist_flat, 0.05f, &w1, &p1, &w2, &p2);
  if (fabs(p2 - p1) < 2 * 0.05f && w2 > 0.5f * w1) {
    w1 += w2;
    p1 = 0.5f * (p1 + p2);
  }
  if (w1 < 150 || p1 < 0.6f) use_flat = 0;
  if (use_flat) {
    float t = 0.9f * p1;
    if (t < 0.1f) t = 0.1f;
    if (t > 0.95f) t = 0.95f;
    s->prior_pars[1] = t;
  }
  two_peaks(s->hist_diff, 0.1f, &w1, &p1, &w2, &p2);
  if (fabs(p2 - p1) < 2 * 0.1f && w2 > 0.5f * w1) {
    w1 += w2;
    p1 = 0.5f * (p1 + p2);
  }
  s->prior_pars[3] = 1.2f * p1;
  if (w1 < 150) use_diff = 0;
  if (s->prior_pars[3] < 0.16f) s->prior_pars[3] = 0.16f;
  if (s->prior_pars[3] > 1.f) s->prior_pars[3] = 1.f;
  if (fluct < 0.05f) use_diff = 0;
  fsum = (float)(1 + use_flat + use_diff);
  s->prior_pars[4] = 1.f / fsum;
  s->prior_pars[5] = (float)use_flat / fsum;
  s->prior_pars[6] = (float)use_diff / fsum;
  memset(s->hist_lrt, 0, sizeof(s->hist_lrt));
  memset(s->hist_flat, 0, sizeof(s->hist_flat));
  memset(s->hist_diff, 0, sizeof(s->hist_diff));
}

void nsf_oracle_analyze(NsfOracle* s, const float* frame) {
  float win[ANA_MAX], re[NB_MAX], im[NB_MAX], magn[NB_MAX], noise[NB_MAX], lmagn[NB_MAX];
  float snr_prior[NB_MAX], snr_post[NB_MAX];
  float energy, sig_e = 0.f, sum_magn = 0.f;
  float sli = 0.f, slisq = 0.f, slm = 0.f, slilm = 0.f;
  const int n = s->nbin, flag = s->upd_flag;
  int i, k, offset = 0;

  push(s->analyze_buf, frame, s->frame, s->ana);
  energy = windowed_energy(s, s->analyze_buf, win);
  if (energy == 0.0) return;
  s->block_ind++;
  spectrum(s, win, re, im, magn);
  for (i = 0; i < n; ++i) {
    sig_e += re[i] * re[i] + im[i] * im[i];
    sum_magn += magn[i];
    if (s->block_ind < 50 && i >= 5) {
      const float li = (float)log((double)(float)i), lm = (float)log((double)magn[i]);
      sli += li;
      slisq += li * li;
      slm += lm;
      slilm += li * lm;
    }
  }
  sig_e = sig_e / (float)n;
  s->signal_energy = sig_e;
  s->sum_magn = sum_magn;

  /* quantile noise tracker */
  if (s->updates < 200) s->updates++;
  for (i = 0; i < n; ++i) lmagn[i] = (float)log((double)magn[i]);
  for (k = 0; k < 3; ++k) {
    offset = k * n;
    for (i = 0; i < n; ++i) {
      float delta = s->density[offset + i] > 1.0 ? 40.f * 1.f / s->density[offset + i] : 40.f;
      if (lmagn[i] > s->lquantile[offset + i]) s->lquantile[offset + i] += 0.25f * delta / (float)(s->counter[k] + 1);
      else s->lquantile[offset + i] -= (1.f - 0.25f) * delta / (float)(s->counter[k] + 1);
      if (fabs(lmagn[i] - s->lquantile[offset + i]) < 0.01f)
        s->density[offset + i] =
            ((float)s->counter[k] * s->density[offset + i] + 1.f / (2.f * 0.01f)) / (float)(s->counter[k] + 1);
    }
    if (s->counter[k] >= 200) {
      s->counter[k] = 0;
      if (s->updates >= 200)
        for (i = 0; i < n; ++i) s->quantile[i] = (float)exp((double)s->lquantile[offset + i]);
    }
    s->counter[k]++;
  }
  if (s->updates < 200)
    for (i = 0; i < n; ++i) s->quantile[i] = (float)exp((double)s->lquantile[offset + i]);
  for (i = 0; i < n; ++i) noise[i] = s->quantile[i];

  if (s->block_ind < 50) {
    float t1, t2, t3, pnum = 0.f, pexp = 0.f;
    s->white_level += sum_magn / (float)n * s->overdrive;
    t1 = slisq * (float)(n - 5);
    t1 -= sli * sli;
    t2 = slisq * slm - sli * slilm;
    t3 = t2 / t1;
    if (t3 < 0.f) t3 = 0.f;
    s->pink_num += t3;
    t2 = sli * slm;
    t2 -= (float)(n - 5) * slilm;
    t3 = t2 / t1;
    if (t3 < 0.f) t3 = 0.f;
    if (t3 > 1.f) t3 = 1.f;
    s->pink_exp += t3;
    if (s->pink_exp > 0.f) {
      pnum = (float)exp((double)(s->pink_num / (float)(s->block_ind + 1)));
      pnum *= (float)(s->block_ind + 1);
      pexp = s->pink_exp / (float)(s->block_ind + 1);
    }
    for (i = 0; i < n; ++i) {
      if (s->pink_exp == 0.f) {
        s->parametric[i] = s->white_level;
      } else {
        const float ub = (float)(i < 5 ? 5 : i);
        s->parametric[i] = (float)(pnum / pow((double)ub, (double)pexp));
      }
      noise[i] *= (float)s->block_ind;
      t2 = s->parametric[i] * (float)(50 - s->block_ind);
      noise[i] += t2 / (float)(s->block_ind + 1);
      noise[i] /= 50.f;
    }
  }
  if (s->block_ind < 200) {
    s->feat[5] *= (float)s->block_ind;
    s->feat[5] += sig_e;
    s->feat[5] /= (float)(s->block_ind + 1);
  }
  /* post / prior SNR */
  for (i = 0; i < n; ++i) {
    const float prev = s->magn_prev_analyze[i] / (s->noise_prev[i] + 0.0001f) * s->smooth[i];
    snr_post[i] = 0.f;
    if (magn[i] > noise[i]) snr_post[i] = magn[i] / (noise[i] + 0.0001f) - 1.f;
    snr_prior[i] = 0.98f * prev + (1.f - 0.98f) * snr_post[i];
  }
  /* spectral flatness */
  {
    float num = 0.f, den = sum_magn - magn[0], sf;
    for (i = 1; i < n; ++i) num += lmagn[i];
    den = den / (float)n;
    num = num / (float)n;
    sf = (float)exp((double)num) / den;
    s->feat[0] += 0.3f * (sf - s->feat[0]);
  }
  /* spectral difference */
  {
    float avg_p = 0.f, avg_m = sum_magn, cov = 0.f, var_p = 0.f, var_m = 0.f, ad;
    for (i = 0; i < n; ++i) avg_p += s->magn_avg_pause[i];
    avg_p = avg_p / (float)n;
    avg_m = avg_m / (float)n;
    for (i = 0; i < n; ++i) {
      cov += (magn[i] - avg_m) * (s->magn_avg_pause[i] - avg_p);
      var_p += (s->magn_avg_pause[i] - avg_p) * (s->magn_avg_pause[i] - avg_p);
      var_m += (magn[i] - avg_m) * (magn[i] - avg_m);
    }
    cov = cov / (float)n;
    var_p = var_p / (float)n;
    var_m = var_m / (float)n;
    s->feat[6] += sig_e;
    ad = var_m - (cov * cov) / (var_p + 0.0001f);
    ad = (float)(ad / (s->feat[5] + 0.0001f));
    s->feat[4] += 0.3f * (ad - s->feat[4]);
  }
  if (flag >= 1) {
    s->upd_count--;
    if (s->upd_count > 0) {
      if (s->feat[3] < 1000 * 0.1f && s->feat[3] >= 0.0) s->hist_lrt[(int)(s->feat[3] / 0.1f)]++;
      if (s->feat[0] < 1000 * 0.05f && s->feat[0] >= 0.0) s->hist_flat[(int)(s->feat[0] / 0.05f)]++;
      if (s->feat[4] < 1000 * 0.1f && s->feat[4] >= 0.0) s->hist_diff[(int)(s->feat[4] / 0.1f)]++;
    }
    if (s->upd_count == 0) {
      extract_thresholds(s);
      s->upd_count = 500;
      if (flag == 1) {
        s->upd_flag = 0;
      } else {
        s->feat[6] = s->feat[6] / 500.f;
        s->feat[5] = 0.5f * (s->feat[6] + s->feat[5]);
        s->feat[6] = 0.f;
      }
    }
  }
  /* speech / noise probability */
  {
    float ksum = 0.f, width, ind0, ind1, ind2, ind_prior, gain_prior, t;
    const float thr0 = s->prior_pars[0], thr1 = s->prior_pars[1], thr2 = s->prior_pars[3];
    const int sgn = (int)s->prior_pars[2];
    for (i = 0; i < n; ++i) {
      const float t1 = 1.f + 2.f * snr_prior[i];
      const float t2 = 2.f * snr_prior[i] / (t1 + 0.0001f);
      const float bessel = (snr_post[i] + 1.f) * t2;
      s->log_lrt[i] += 0.5f * (bessel - (float)log((double)t1) - s->log_lrt[i]);
      ksum += s->log_lrt[i];
    }
    ksum = ksum / (float)n;
    s->feat[3] = ksum;
    width = ksum < thr0 ? 8.f : 4.f;
    ind0 = 0.5f * ((float)tanh((double)(width * (ksum - thr0))) + 1.f);
    t = s->feat[0];
    width = 4.f;
    if (sgn == 1 && t > thr1) width = 8.f;
    if (sgn == -1 && t < thr1) width = 8.f;
    ind1 = 0.5f * ((float)tanh((double)((float)sgn * width * (thr1 - t))) + 1.f);
    t = s->feat[4];
    width = t < thr2 ? 8.f : 4.f;
    ind2 = 0.5f * ((float)tanh((double)(width * (t - thr2))) + 1.f);
    ind_prior = s->prior_pars[4] * ind0 + s->prior_pars[5] * ind1 + s->prior_pars[6] * ind2;
    s->prior_speech_prob += 0.1f * (ind_prior - s->prior_speech_prob);
    if (s->prior_speech_prob > 1.f) s->prior_speech_prob = 1.f;
    if (s->prior_speech_prob < 0.01f) s->prior_speech_prob = 0.01f;
    gain_prior = (1.f - s->prior_speech_prob) / (s->prior_speech_prob + 0.0001f);
    for (i = 0; i < n; ++i) {
      float inv = (float)exp((double)-s->log_lrt[i]);
      inv = gain_prior * inv;
      s->speech_prob[i] = 1.f / (1.f + inv);
    }
  }
  /* noise update, gamma carried from the previous bin */
  {
    float gamma = 0.9f;
    for (i = 0; i < n; ++i) {
      const float ps = s->speech_prob[i], pn = 1.f - ps;
      const float tmp = gamma * s->noise_prev[i] + (1.f - gamma) * (pn * magn[i] + ps * s->noise_prev[i]);
      const float old = gamma;
      gamma = ps > 0.2f ? 0.99f : 0.9f;
      if (ps < 0.2f) s->magn_avg_pause[i] += 0.05f * (magn[i] - s->magn_avg_pause[i]);
      if (gamma == old) {
        noise[i] = tmp;
      } else {
        noise[i] = gamma * s->noise_prev[i] + (1.f - gamma) * (pn * magn[i] + ps * s->noise_prev[i]);
        if (tmp < noise[i]) noise[i] = tmp;
      }
    }
  }
  memcpy(s->noise, noise, sizeof(float) * (size_t)n);
  memcpy(s->magn_prev_analyze, magn, sizeof(float) * (size_t)n);
}

static float sat(float v) { return v > 32767.f ? 32767.f : (v < -32768.f ? -32768.f : v); }

void nsf_oracle_process(NsfOracle* s, const float* const* in, int num_bands, float* const* out) {
  float win[ANA_MAX], re[NB_MAX], im[NB_MAX], magn[NB_MAX], filt[NB_MAX];
  float energy1, factor = 1.f;
  const int n = s->nbin, nhb = num_bands - 1;
  int i, b;

  push(s->data_buf, in[0], s->frame, s->ana);
  for (b = 0; b < nhb; ++b) push(s->hb_buf[b], in[b + 1], s->frame, s->ana);
  energy1 = windowed_energy(s, s->data_buf, win);
  if (energy1 == 0.0) {
    for (i = 0; i < s->frame; ++i) out[0][i] = sat(s->synt_buf[i]);
    push(s->synt_buf, NULL, s->frame, s->ana);
    for (b = 0; b < nhb; ++b)
      for (i = 0; i < s->frame; ++i) out[b + 1][i] = sat(s->hb_buf[b][i]);
    return;
  }
  spectrum(s, win, re, im, magn);
  if (s->block_ind < 50)
    for (i = 0; i < n; ++i) s->init_magn[i] += magn[i];
  for (i = 0; i < n; ++i) {
    const float prev = s->magn_prev_process[i] / (s->noise_prev[i] + 0.0001f) * s->smooth[i];
    float cur = 0.f, prior;
    if (magn[i] > s->noise[i]) cur = magn[i] / (s->noise[i] + 0.0001f) - 1.f;
    prior = 0.98f * prev + (1.f - 0.98f) * cur;
    filt[i] = prior / (s->overdrive + prior);
  }
  for (i = 0; i < n; ++i) {
    if (filt[i] < s->denoise_bound) filt[i] = s->denoise_bound;
    if (filt[i] > 1.f) filt[i] = 1.f;
    if (s->block_ind < 50) {
      float ft = s->init_magn[i] - s->overdrive * s->parametric[i];
      ft /= (s->init_magn[i] + 0.0001f);
      if (ft < s->denoise_bound) ft = s->denoise_bound;
      if (ft > 1.f) ft = 1.f;
      filt[i] *= (float)s->block_ind;
      ft *= (float)(50 - s->block_ind);
      filt[i] += ft;
      filt[i] /= 50.f;
    }
    s->smooth[i] = filt[i];
    re[i] *= s->smooth[i];
    im[i] *= s->smooth[i];
  }
  memcpy(s->magn_prev_process, magn, sizeof(float) * (size_t)n);
  memcpy(s->noise_prev, s->noise, sizeof(float) * (size_t)n);
  rfft_inv(s, re, im, win);
  if (s->gainmap == 1 && s->block_ind > 200) {
    float factor1 = 1.f, factor2 = 1.f, energy2 = 0.f, gain;
    for (i = 0; i < s->ana; ++i) energy2 += win[i] * win[i];
    gain = (float)sqrt((double)(energy2 / (energy1 + 1.f)));
    if (gain > 0.5f) {
      factor1 = 1.f + 1.3f * (gain - 0.5f);
      if (gain * factor1 > 1.f) factor1 = 1.f / gain;
    }
    if (gain < 0.5f) {
      if (gain <= s->denoise_bound) gain = s->denoise_bound;
      factor2 = 1.f - 0.3f * (0.5f - gain);
    }
    factor = s->prior_speech_prob * factor1 + (1.f - s->prior_speech_prob) * factor2;
  }
  for (i = 0; i < s->ana; ++i) win[i] = s->window[i] * win[i];
  for (i = 0; i < s->ana; ++i) s->synt_buf[i] += factor * win[i];
  for (i = 0; i < s->frame; ++i) out[0][i] = sat(s->synt_buf[i]);
  push(s->synt_buf, NULL, s->frame, s->ana);

  if (nhb > 0) {
    const int d = n / 4;
    float avg_prob = 0.f, avg_gain = 0.f, sa = 0.f, sp = 0.f, gmod, g;
    for (i = n - d - 1; i < n - 1; ++i) avg_prob += s->speech_prob[i];
    avg_prob = avg_prob / (float)d;
    for (i = 0; i < n; ++i) {
      sa += s->magn_prev_analyze[i];
      sp += s->magn_prev_process[i];
    }
    avg_prob *= sp / sa;
    for (i = n - d - 1; i < n - 1; ++i) avg_gain += s->smooth[i];
    avg_gain = avg_gain / (float)d;
    gmod = 0.5f * (1.f + (float)tanh((double)(2.f * avg_prob - 1.f)));
    g = 0.5f * gmod + 0.5f * avg_gain;
    if (avg_prob >= 0.5f) g = 0.25f * gmod + 0.75f * avg_gain;
    if (g < s->denoise_bound) g = s->denoise_bound;
    if (g > 1.f) g = 1.f;
    for (b = 0; b < nhb; ++b)
      for (i = 0; i < s->frame; ++i) out[b + 1][i] = sat(g * s->hb_buf[b][i]);
  }
}

int nsf_oracle_run(int fs, int mode, int nframes, const int16_t* pcm_in, float* out_f32, float* prior_prob) {
  NsfOracle* s = nsf_oracle_create();
  float in[160], out[160];
  int f, i;
  if (!s || fs > 16000 || nsf_oracle_init(s, (uint32_t)fs) != 0 || nsf_oracle_set_policy(s, mode) != 0) {
    nsf_oracle_free(s);
    return -1;
  }
  for (f = 0; f < nframes; ++f) {
    const float* ib[1] = {in};
    float* ob[1] = {out};
    for (i = 0; i < s->frame; ++i) in[i] = (float)pcm_in[(size_t)f * s->frame + i];
    nsf_oracle_analyze(s, in);
    nsf_oracle_process(s, ib, 1, ob);
    memcpy(out_f32 + (size_t)f * s->frame, out, sizeof(float) * (size_t)s->frame);
    if (prior_prob) prior_prob[f] = s->prior_speech_prob;
  }
  nsf_oracle_free(s);
  return 0;
}
