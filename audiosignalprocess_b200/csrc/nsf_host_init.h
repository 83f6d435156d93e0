// Host-side construction of the float-NS constant tables and of a freshly
// initialised per-stream state slab (the job of WebRtcNs_InitCore,
// ns_core.c:74-214, and WebRtcNs_set_policy_core, ns_core.c:1013-1041).
// Plain C++; no CUDA calls.  Used by the C-ABI host layer (ns_capi.cu) and by
// the test-only SIMT emulator.
#ifndef AUDIOSIGNALPROCESS_B200_NSF_HOST_INIT_H_
#define AUDIOSIGNALPROCESS_B200_NSF_HOST_INIT_H_

#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector_functions.h>

#include "nsf_layout.h"

namespace nsb200 {

// Hybrid Hann/flat window (windows_private.h:64,94): a sine rise over the
// overlap, flat top, mirrored fall.  The reference tabulates it to 8 decimals;
// rounding sin() to 8 decimals in double and then to float reproduces the
// table bit for bit (checked in tests/test_tables.py against the reference).
inline void nsf_make_window(int ana, float* w) {
  const int rise = ana == 256 ? 96 : 48;
  const double kPi = 3.14159265358979323846;
  for (int i = 0; i < ana; ++i) {
    double v;
    if (i < rise) v = sin(kPi * i / (2.0 * rise));
    else if (i <= ana - rise) v = 1.0;
    else v = sin(kPi * (ana - i) / (2.0 * rise));
    w[i] = (float)(floor(v * 1e8 + 0.5) / 1e8);
  }
}

// ---- tables of the forward FFT, which reproduces the rounding order of the reference's Ooura rdft
// (ns_warp.cuh ooura_fwd; utility/fft4g.c).  Everything below is evaluated in the reference's own
// precision and order -- angles formed in float, cos/sin in double rounded to float, the third
// twiddle of a group derived from the other two in float -- because the kernel must multiply by the
// very same floats (checked against WebRtc_rdft through the emulator in tests/test_emulated_kernels.py
// and on the device in tests/test_gpu_float_parity.py).
inline unsigned nsf_bit_reverse(unsigned v, int bits) {
  unsigned r = 0;
  for (int b = 0; b < bits; ++b) r |= ((v >> b) & 1u) << (bits - 1 - b);
  return r;
}
// Quarter-circle twiddles W[q] = (cos, sin)(pi q / (n/4)), q < n/8, in bit-reversed order (makewt, fft4g.c:642).
inline void nsf_ooura_quarter(int n, float2* W) {
  const int nw = n >> 2, nwh = nw >> 1, nq = nw >> 1;
  int qbits = 0;
  while ((1 << qbits) < nq) ++qbits;
  const float delta = (float)atan(1.0) / (float)nwh;
  float w[64];
  w[0] = 1.f;
  w[1] = 0.f;
  w[nwh] = (float)cos((double)(delta * (float)nwh));
  w[nwh + 1] = w[nwh];
  for (int j = 2; j < nwh; j += 2) {
    const float ang = delta * (float)j;
    const float x = (float)cos((double)ang), y = (float)sin((double)ang);
    w[j] = x;
    w[j + 1] = y;
    w[nw - j] = y;
    w[nw - j + 1] = x;
  }
  for (int q = 0; q < nq; ++q) {
    const unsigned r = nsf_bit_reverse((unsigned)q, qbits);
    W[q] = make_float2(w[2 * r], w[2 * r + 1]);
  }
}
// The three output twiddles of butterfly group g of any radix-4 pass (cft1st / cftmdl, fft4g.c:1002,1107).
// Group 1 is the pi/4 group, evaluated by the reference as c * (a -+ b): stored as (c, 0), i, (-c, 0) for the
// kernel's `diag` form.  volatile: the float products must round exactly once, whatever the host compiler does.
inline void nsf_ooura_group(const float2* W, int g, float2* w1, float2* w2, float2* w3) {
  *w1 = *w2 = *w3 = make_float2(1.f, 0.f);
  if (g == 0) return;
  const int p = g >> 1;
  const float ar = W[p].x, ai = W[p].y;
  const bool odd = (g & 1) != 0;
  *w1 = odd ? W[2 * p + 1] : W[2 * p];
  *w2 = odd ? make_float2(-ai, ar) : make_float2(ar, ai);
  const float h = odd ? ar : ai;
  volatile float t = 2.f * h;
  volatile float u = t * w1->y;
  volatile float r3 = w1->x - u;
  u = t * w1->x;
  volatile float i3 = u - w1->y;
  *w3 = make_float2(r3, i3);
  if (g == 1) {
    *w1 = make_float2(w1->x, 0.f);
    *w3 = make_float2(-w1->x, 0.f);
  }
}
// img_split: 129 float2 (real-input split per bin, rftfsub fft4g.c:1234 + :352-354); img_otw: kOouraTwF2 float2.
inline void nsf_fill_ooura(int ana, float2* img_split, float2* img_otw) {
  const int nc = ana / 2, ncq = ana >> 2, nch = ncq >> 1;
  float2 W[32];
  nsf_ooura_quarter(ana, W);
  // half-cosine table of makect (fft4g.c:671)
  float c[65];
  {
    const float delta = (float)atan(1.0) / (float)nch;
    c[0] = (float)cos((double)(delta * (float)nch));
    c[nch] = 0.5f * c[0];
    for (int j = 1; j < nch; ++j) {
      const float ang = delta * (float)j;
      c[j] = 0.5f * (float)cos((double)ang);
      c[ncq - j] = 0.5f * (float)sin((double)ang);
    }
  }
  for (int k = 0; k <= nc; ++k) {
    float2 w = make_float2(0.f, 0.f);            // bin nc/2 passes through
    if (k == 0) w = make_float2(0.f, 0.5f);       // a[0] + a[1]
    else if (k == nc) w = make_float2(0.f, -0.5f);  // a[0] - a[1]
    else if (k < nc / 2) w = make_float2(0.5f - c[ncq - k], c[k]);
    else if (k > nc / 2) w = make_float2(0.5f - c[ncq - (nc - k)], -c[nc - k]);
    img_split[k] = w;
  }
  // pass 1: lane H owns group rev(H); pass 2: index l owns group rev(l)
  const int lanes = nc / 4, lbits = lanes == 32 ? 5 : 4, gbits = lanes == 32 ? 3 : 2;
  for (int H = 0; H < lanes; ++H)
    nsf_ooura_group(W, (int)nsf_bit_reverse((unsigned)H, lbits), &img_otw[H], &img_otw[32 + H], &img_otw[64 + H]);
  for (int l = 0; l < (1 << gbits); ++l)
    nsf_ooura_group(W, (int)nsf_bit_reverse((unsigned)l, gbits), &img_otw[96 + l], &img_otw[96 + 8 + l], &img_otw[96 + 16 + l]);
  img_otw[120] = make_float2(W[1].x, -W[1].x);   // pass 3, group 1: c = cos(pi/4) as the reference rounds it
}

template <typename Tables>
inline void nsf_fill_tables(Tables* t) {
  memset(t, 0, sizeof(*t));
  nsf_make_window(256, t->win256);
  nsf_make_window(128, t->win128);
  const double kPi = 3.14159265358979323846;
  for (int k = 0; k < 256; ++k) {
    t->tw[k].x = (float)cos(2.0 * kPi * k / 256.0);
    t->tw[k].y = (float)sin(2.0 * kPi * k / 256.0);
  }
  t->tw[0].x = 1.f;   t->tw[0].y = 0.f;
  t->tw[64].x = 0.f;  t->tw[64].y = 1.f;
  t->tw[128].x = -1.f; t->tw[128].y = 0.f;
  t->tw[192].x = 0.f; t->tw[192].y = -1.f;
  // the eighth roots the warp FFT's last pass writes out by hand (ns_warp.cuh): one constant
  const float c8 = (float)cos(kPi / 4.0);
  t->tw[32].x = c8;   t->tw[32].y = c8;
  t->tw[96].x = -c8;  t->tw[96].y = c8;
  t->tw[160].x = -c8; t->tw[160].y = -c8;
  t->tw[224].x = c8;  t->tw[224].y = -c8;
  for (int i = 1; i < 132; ++i) t->logi[i] = (float)log((double)(float)i);
  for (int i = 1; i < 132; ++i) t->logk_d[i] = log((double)i);
  // sequential float sums exactly as ns_core.c:1088-1100 accumulates them
  for (int v = 0; v < 2; ++v) {
    const int magn_len = v == 0 ? 129 : 65;
    float s = 0.f, s2 = 0.f;
    for (int i = 5; i < magn_len; ++i) {
      const float l = t->logi[i];
      s += l;
      s2 += l * l;
    }
    t->sum_log_i[v] = s;
    t->sum_log_i_sq[v] = s2;
  }
  // shared-memory images (nsf_layout.h); the regrouped twiddles follow ns_warp.cuh fft_fill_tw12
  for (int v = 0; v < 2; ++v) {
    float* img = t->img[v];
    const int ana = v == 0 ? 256 : 128, nc = ana / 2, l = nc / 4, m = l / 4;
    memcpy(img + kNsfImgWin, v == 0 ? t->win256 : t->win128, sizeof(float) * ana);
    memcpy(img + kNsfImgTw, t->tw, sizeof(t->tw));
    memcpy(img + kNsfImgLogi, t->logi, sizeof(t->logi));
    float2* tw12 = reinterpret_cast<float2*>(img + kNsfImgTw12);
    for (int i = 0; i < 96; ++i) tw12[i] = t->tw[((i & 31) * (i / 32 + 1) * (256 / nc)) & 255];
    for (int i = 0; i < 24; ++i) tw12[96 + i] = t->tw[(((i & 7) % m) * (i / 8 + 1) * (256 / l)) & 255];
    nsf_fill_ooura(ana, reinterpret_cast<float2*>(img + kNsfImgSplit), reinterpret_cast<float2*>(img + kNsfImgOtw));
  }
}

inline bool nsf_mode_params(int mode, float* overdrive, float* denoise_bound, int* gainmap) {
  switch (mode) {  // ns_core.c:1020-1039
    case 0: *overdrive = 1.f;   *denoise_bound = 0.5f;   *gainmap = 0; return true;
    case 1: *overdrive = 1.f;   *denoise_bound = 0.25f;  *gainmap = 1; return true;
    case 2: *overdrive = 1.1f;  *denoise_bound = 0.125f; *gainmap = 1; return true;
    case 3: *overdrive = 1.25f; *denoise_bound = 0.09f;  *gainmap = 1; return true;
  }
  return false;
}

inline void nsf_set_mode(uint32_t* slab, int mode) {
  float od = 1.f, db = 0.5f;
  int gm = 0;
  nsf_mode_params(mode, &od, &db, &gm);
  memcpy(slab + kH_overdrive, &od, 4);
  memcpy(slab + kH_denoiseBound, &db, 4);
  slab[kH_gainmap] = (uint32_t)gm;
  slab[kH_mode] = (uint32_t)mode;
}

inline void nsf_init_state(uint32_t* slab, uint32_t fs) {
  memset(slab, 0, sizeof(uint32_t) * (size_t)kNsfStateWords);
  float* f = reinterpret_cast<float*>(slab);
  int32_t* i = reinterpret_cast<int32_t*>(slab);
  i[kH_blockInd] = -1;
  i[kH_updates] = 0;
  for (int s = 0; s < 3; ++s)
    i[kH_counter + s] = (int)floor((float)(200 * (s + 1)) / 3.f);  // 66, 133, 200
  i[kH_modelUpd0] = 2;
  i[kH_modelUpd3] = 500;
  i[kH_splitValid] = 1;   // all-zero split arrays are what InitCore leaves (ns_core.c:103-131)
  const float pars[7] = {0.5f, 0.5f, 1.f, 0.5f, 1.f, 0.f, 0.f};
  for (int k = 0; k < 7; ++k) f[kH_priorPars + k] = pars[k];
  f[kH_priorSpeechProb] = 0.5f;
  const float feat[7] = {0.5f, 0.f, 0.f, 0.5f, 0.5f, 0.f, 0.f};
  for (int k = 0; k < 7; ++k) f[kH_feat + k] = feat[k];
  i[kH_fs] = (int32_t)fs;
  i[kH_initFlag] = 1;
  for (int b = 0; b < 129; ++b) {
    float* r = f + kNsfOffBins + b * kNsfBinRec;
    for (int s = 0; s < 3; ++s) {
      r[kB_lq0 + s] = 8.f;
      r[kB_dens0 + s] = 0.3f;
    }
    r[kB_smooth] = 1.f;
    r[kB_logLrt] = 0.5f;
  }
  nsf_set_mode(slab, 0);
}

}  // namespace nsb200

#endif  // AUDIOSIGNALPROCESS_B200_NSF_HOST_INIT_H_
