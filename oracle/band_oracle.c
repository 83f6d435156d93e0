/* TEST INFRASTRUCTURE ONLY -- never linked into or called from the product.
 *
 * CPU restatement of the 32/48 kHz band split / merge that AudioBuffer wraps
 * around the suppressors (root WebRtc_AMP_Port/webrtc/):
 *   WebRtcSpl_AllPassQMF / AnalysisQMF / SynthesisQMF
 *       common_audio/signal_processing/splitting_filter_c.c:48,127,167
 *   SplittingFilter::TwoBands* / ThreeBands*
 *       modules/audio_processing/splitting_filter.cc:65-171
 *   PushSincResampler::Resample (int16), SincResampler::Resample /
 *   InitializeKernel, Convolve_SSE
 *       common_audio/resampler/push_sinc_resampler.cc:33-100,
 *       sinc_resampler.cc:151-242,269-342, sinc_resampler_sse.cc:20-57
 * Integer parts are bit-exact.  The resampler keeps the x86 reference's float
 * operation order (four interleaved partial sums, interpolation, (s0+s2)+(s1+s3))
 * and its running double position, so the int16 results are bit-exact as well
 * (pinned in tests/test_oracle_pinning.py against oracle/_ref).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ns_oracle.h"
#include "spl_fixed.h"

/* ---- all-pass QMF ------------------------------------------------------------ */
static const uint16_t kAp1[3] = {6418, 36982, 57261};
static const uint16_t kAp2[3] = {21333, 49062, 63010};

static int32_t sub_sat32(int32_t a, int32_t b) {
  int64_t d = (int64_t)a - (int64_t)b;
  return d > 2147483647LL ? 2147483647 : (d < -2147483648LL ? (int32_t)0x80000000 : (int32_t)d);
}
static int32_t scale_diff(uint16_t a, int32_t b, int32_t c) {
  return (int32_t)((uint32_t)c + (uint32_t)(b >> 16) * a + ((((uint32_t)b & 0xFFFFu) * a) >> 16));
}
/* three cascaded first-order sections, sample by sample; state as in the reference:
 * {x[-1], y1[-1], y1[-1], y2[-1], y2[-1], y3[-1]} */
static void allpass(const int32_t* in, int n, int32_t* out, const uint16_t* c, int32_t* st) {
  int k;
  for (k = 0; k < n; ++k) {
    const int32_t x = in[k];
    const int32_t y1 = scale_diff(c[0], sub_sat32(x, st[1]), st[0]);
    const int32_t y2 = scale_diff(c[1], sub_sat32(y1, st[3]), st[2]);
    const int32_t y3 = scale_diff(c[2], sub_sat32(y2, st[5]), st[4]);
    st[0] = x; st[1] = y1; st[2] = y1; st[3] = y2; st[4] = y2; st[5] = y3;
    out[k] = y3;
  }
}

void band_oracle_qmf_analysis(const int16_t* in, int len, int16_t* low, int16_t* high, int32_t* st1,
                              int32_t* st2) {
  int32_t a[320] = {0}, b[320] = {0}, fa[320], fb[320];
  const int h = len / 2;
  int i;
  for (i = 0; i < h; ++i) {
    b[i] = (int32_t)in[2 * i] * 1024;
    a[i] = (int32_t)in[2 * i + 1] * 1024;
  }
  allpass(a, h, fa, kAp1, st1);
  allpass(b, h, fb, kAp2, st2);
  for (i = 0; i < h; ++i) {
    low[i] = fx_sat16((fa[i] + fb[i] + 1024) >> 11);
    high[i] = fx_sat16((fa[i] - fb[i] + 1024) >> 11);
  }
}

void band_oracle_qmf_synthesis(const int16_t* low, const int16_t* high, int band_len, int16_t* out,
                               int32_t* st1, int32_t* st2) {
  int32_t a[320] = {0}, b[320] = {0}, fa[320], fb[320];
  int i;
  for (i = 0; i < band_len; ++i) {
    a[i] = ((int32_t)low[i] + (int32_t)high[i]) * 1024;
    b[i] = ((int32_t)low[i] - (int32_t)high[i]) * 1024;
  }
  allpass(a, band_len, fa, kAp2, st1);
  allpass(b, band_len, fb, kAp1, st2);
  for (i = 0; i < band_len; ++i) {
    out[2 * i] = fx_sat16((fb[i] + 512) >> 10);
    out[2 * i + 1] = fx_sat16((fa[i] + 512) >> 10);
  }
}

/* ---- push sinc resampler ------------------------------------------------------ */
typedef struct {
  int src, dst, block, primed, first_pass, second_load;
  double ratio, vsi;
  float kernel[33 * 32];
  float buf[640 + 32];
} Resampler;

static void resampler_init(Resampler* r, int src, int dst) {
  const double pi = 3.14159265358979323846;
  const double a0 = 0.5 * (1.0 - 0.16), a1 = 0.5, a2 = 0.5 * 0.16;
  double scale;
  int o, i;
  memset(r, 0, sizeof(*r));
  r->src = src;
  r->dst = dst;
  r->ratio = src * 1.0 / dst;
  r->first_pass = 1;
  r->block = src - 16;                        /* r0 = buffer + 16 until the first refill */
  scale = (r->ratio > 1.0 ? 1.0 / r->ratio : 1.0) * 0.9;
  for (o = 0; o <= 32; ++o) {
    const float sub = (float)o / 32;
    for (i = 0; i < 32; ++i) {
      const float pre = (float)(pi * (i - 16 - sub));
      const float x = (i - sub) / 32;
      const float w = (float)(a0 - a1 * cos(2.0 * pi * x) + a2 * cos(4.0 * pi * x));
      r->kernel[o * 32 + i] = (float)(w * ((pre == 0) ? scale : (sin(scale * pre) / pre)));
    }
  }
}

static float convolve(const float* in, const float* k1, const float* k2, double factor) {
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, t[4];
  const float f1 = (float)(1.0 - factor), f2 = (float)factor;
  int i, c;
  for (i = 0; i < 32; i += 4)
    for (c = 0; c < 4; ++c) {
      s1[c] += in[i + c] * k1[i + c];
      s2[c] += in[i + c] * k2[i + c];
    }
  for (c = 0; c < 4; ++c) t[c] = s1[c] * f1 + s2[c] * f2;
  return (t[0] + t[2]) + (t[1] + t[3]);
}

/* SincResampler::Resample with the PushSincResampler callback folded in: `frame`
 * is consumed when the current block runs out (NULL = the zero priming block). */
static void sinc_run(Resampler* r, int frames, float* dest, const int16_t* frame) {
  float* r0 = r->buf + (r->second_load ? 32 : 16);
  int remaining = frames, i, k;
  if (!r->primed && remaining) {
    for (k = 0; k < r->src; ++k) r0[k] = frame ? (float)frame[k] : 0.f;
    frame = NULL;
    r->primed = 1;
  }
  while (remaining) {
    for (i = (int)ceil((r->block - r->vsi) / r->ratio); i > 0; --i) {
      const int sidx = (int)r->vsi;
      const double voff = (r->vsi - sidx) * 32;
      const int off = (int)voff;
      *dest++ = convolve(r->buf + sidx, r->kernel + off * 32, r->kernel + off * 32 + 32, voff - off);
      r->vsi += r->ratio;
      if (!--remaining) return;
    }
    r->vsi -= r->block;
    memcpy(r->buf, r0 + r->src - 32, sizeof(float) * 32);
    if (!r->second_load) {
      r->second_load = 1;
      r0 = r->buf + 32;
      r->block = r->src;
    }
    for (k = 0; k < r->src; ++k) r0[k] = frame ? (float)frame[k] : 0.f;
    frame = NULL;
  }
}

static void resample_frame(Resampler* r, const int16_t* in, int16_t* out) {
  float tmp[640];
  int i;
  if (r->first_pass) {
    sinc_run(r, (int)(r->block / r->ratio), tmp, NULL);   /* zero-primed pass, output discarded */
    r->first_pass = 0;
  }
  sinc_run(r, r->dst, tmp, in);
  for (i = 0; i < r->dst; ++i) {
    const float v = tmp[i];
    out[i] = v > 0 ? (v >= 32766.5f ? 32767 : (int16_t)(v + 0.5f)) : (v <= -32767.5f ? -32768 : (int16_t)(v - 0.5f));
  }
}

/* ---- splitter ------------------------------------------------------------------- */
struct BandOracle {
  int fs, nb;
  int32_t st[3][4][6];   /* instance x {ana1, ana2, syn1, syn2} */
  Resampler up, down;
};

BandOracle* band_oracle_create(int fs) {
  BandOracle* b = (BandOracle*)calloc(1, sizeof(BandOracle));
  if (!b) return NULL;
  b->fs = fs;
  b->nb = fs == 32000 ? 2 : (fs == 48000 ? 3 : 1);
  resampler_init(&b->up, 480, 640);
  resampler_init(&b->down, 640, 480);
  return b;
}
void band_oracle_free(BandOracle* b) { free(b); }
int band_oracle_num_bands(const BandOracle* b) { return b->nb; }

/* in: fs/100 samples; bands: [nb][160] */
void band_oracle_split(BandOracle* b, const int16_t* in, int16_t* bands) {
  if (b->nb == 1) {
    memcpy(bands, in, sizeof(int16_t) * (size_t)(b->fs / 100));
  } else if (b->nb == 2) {
    band_oracle_qmf_analysis(in, 320, bands, bands + 160, b->st[0][0], b->st[0][1]);
  } else {
    int16_t s64[640], lo[320], hi[320], drop[160];
    resample_frame(&b->up, in, s64);
    band_oracle_qmf_analysis(s64, 640, lo, hi, b->st[0][0], b->st[0][1]);
    band_oracle_qmf_analysis(lo, 320, bands, bands + 160, b->st[1][0], b->st[1][1]);
    band_oracle_qmf_analysis(hi, 320, drop, bands + 320, b->st[2][0], b->st[2][1]);
  }
}

void band_oracle_merge(BandOracle* b, const int16_t* bands, int16_t* out) {
  if (b->nb == 1) {
    memcpy(out, bands, sizeof(int16_t) * (size_t)(b->fs / 100));
  } else if (b->nb == 2) {
    band_oracle_qmf_synthesis(bands, bands + 160, 160, out, b->st[0][2], b->st[0][3]);
  } else {
    int16_t lo[320], hi[320], s64[640], zeros[160];
    memset(zeros, 0, sizeof(zeros));
    band_oracle_qmf_synthesis(bands, bands + 160, 160, lo, b->st[1][2], b->st[1][3]);
    band_oracle_qmf_synthesis(zeros, bands + 320, 160, hi, b->st[2][2], b->st[2][3]);
    band_oracle_qmf_synthesis(lo, hi, 320, s64, b->st[0][2], b->st[0][3]);
    resample_frame(&b->down, s64, out);
  }
}

/* Whole path, one stream: split -> NSx (fixed=1) or float NS -> merge, as
 * AudioBuffer + the suppressor do per 10 ms frame (libapm/src/apm_ns.cpp:96-132). */
int band_oracle_run(int fixed, int fs, int mode, int nframes, const int16_t* pcm_in, int16_t* pcm_out) {
  BandOracle* b = band_oracle_create(fs);
  NsxOracle* sx = NULL;
  NsfOracle* sf = NULL;
  const int n = fs / 100;
  int f, k, i, rc = 0;
  if (!b) return -1;
  if (fixed) {
    sx = nsx_oracle_create();
    rc = (!sx || nsx_oracle_init(sx, (uint32_t)fs) || nsx_oracle_set_policy(sx, mode)) ? -1 : 0;
  } else {
    sf = nsf_oracle_create();
    rc = (!sf || nsf_oracle_init(sf, (uint32_t)fs) || nsf_oracle_set_policy(sf, mode)) ? -1 : 0;
  }
  for (f = 0; f < nframes && rc == 0; ++f) {
    int16_t bands[3 * 160], obands[3 * 160];
    const int bl = b->nb == 1 ? n : 160;
    band_oracle_split(b, pcm_in + (size_t)f * n, bands);
    if (fixed) {
      const int16_t* ib[3] = {bands, bands + 160, bands + 320};
      int16_t* ob[3] = {obands, obands + 160, obands + 320};
      nsx_oracle_process(sx, ib, b->nb, ob);
    } else {
      float fin[3][160], fout[3][160];
      const float* ib[3] = {fin[0], fin[1], fin[2]};
      float* ob[3] = {fout[0], fout[1], fout[2]};
      for (k = 0; k < b->nb; ++k)
        for (i = 0; i < bl; ++i) fin[k][i] = (float)bands[k * 160 + i];
      nsf_oracle_analyze(sf, fin[0]);
      nsf_oracle_process(sf, ib, b->nb, ob);
      for (k = 0; k < b->nb; ++k)
        for (i = 0; i < bl; ++i) {
          const float v = fout[k][i];   /* IFChannelBuffer::RefreshI, channel_buffer.cc:55-60 */
          obands[k * 160 + i] = v > 0 ? (v >= 32766.5f ? 32767 : (int16_t)(v + 0.5f))
                                      : (v <= -32767.5f ? -32768 : (int16_t)(v - 0.5f));
        }
    }
    band_oracle_merge(b, obands, pcm_out + (size_t)f * n);
  }
  nsx_oracle_free(sx);
  nsf_oracle_free(sf);
  band_oracle_free(b);
  return rc;
}
