// Fixed-point noise suppressor (WebRtcNsx_Process) as one sm_100a kernel: one
// warp per stream walks F 10-ms frames; lanes own time samples and bins.  The
// int16 output is bit-identical to the reference.
//
// What it replaces (webrtc/modules/audio_processing/ns/):
//   WebRtcNsx_ProcessCore nsx_core.c:1502-2121, WebRtcNsx_DataAnalysis :1183,
//   WebRtcNsx_DataSynthesis :1421, AnalysisUpdateC :523, NormalizeRealBufferC
//   :543, NoiseEstimationC :334 (+UpdateNoiseEstimate :303), PrepareSpectrumC
//   :455, DenormalizeC :476, SynthesisUpdateC :490, ComputeSpectralFlatness
//   :1021, ComputeSpectralDifference :1090, FeatureParameterExtraction :820,
//   CalcParametricNoiseEstimate :585, WebRtcNsx_SpeechNoiseProb
//   nsx_core_c.c:26, and the SPL helpers they call (ns_fixed.cuh).
//
// Mapping: LANES = ANA/8 active lanes (32 at 16 kHz, 16 at 8 kHz).
//   time sample i  = lane + LANES*r, r = 0..7   (history r<3, frame r<5 of the output)
//   bin k          = lane + LANES*j, j = 0..3; slot 4 = Nyquist bin (lane 0, mirrored to all)
//   FFT position p: three layouts, see ns_fixed.cuh; the forward transform ends
//   in layout C, which IS the bin mapping, so no exchange is needed after it.
// Sums that the reference accumulates in uint32/int32 wrap-around arithmetic are
// order independent, so warp tree reductions are exact.
#ifndef AUDIOSIGNALPROCESS_B200_NSX_KERNEL_CUH_
#define AUDIOSIGNALPROCESS_B200_NSX_KERNEL_CUH_

#include "ns_fixed.cuh"
#include "nsx_layout.h"

namespace nsb200 {

// CTA shape: the loop body is ~7000 instructions (two fully unrolled 256-point int16 FFTs, ~30
// inlined integer divisions), far more than the instruction cache holds once 28 warps per SM run
// in 28 different places: "no instruction" was the dominant stall (profiles/r1_nsx_kernel_F100.md).
// The warps of a CTA therefore meet at a barrier once per frame and walk the code together, and
// the host picks the CTA size at launch (blockDim) so that a batch spreads over all SMs with
// as many lock-stepped warps per CTA as possible: one 28-warp CTA per SM at 4096 streams
// (measured 4.07 -> 2.99 ms per 4096 x 100 frames; more barriers per frame were slower).
constexpr int kNsxWarpsPerCta = 2;        // smallest CTA (and the emulator's)
#ifndef NSX_MAX_WARPS
#define NSX_MAX_WARPS 28
#endif
constexpr int kNsxMaxWarpsPerCta = NSX_MAX_WARPS;    // 72 registers x 896 threads, 196 KB of shared memory
#ifndef NSX_FRAME_SYNC
#define NSX_FRAME_SYNC 1
#endif
// per CTA: table image (window | twiddles | log2 fraction table) | mbarriers of the TMA bulk copies
// (one for the tables, two per warp: header + sample histories | per-bin records)
constexpr int kNsxCtaBarWords = (2 * (1 + 2 * kNsxMaxWarpsPerCta) + 3) / 4 * 4;
constexpr int kNsxCtaTableWords = kNsxTableImgWords + kNsxCtaBarWords;
constexpr int kNsxScratchWords = 512 + 136;          // FFT transposes (64-bit points) | time / spectrum buffer
constexpr int kNsxWarpWords = 2 * kNsxHdrWords + 2 * 129 * 4 + kNsxScratchWords;

NSB_DEV unsigned warp_sum_u(unsigned v) {
#ifdef __CUDA_ARCH__
  return __reduce_add_sync(kFullMask, v);   // (ns_warp.cuh: integer warp reductions are one REDUX)
#else
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
#endif
}
NSB_DEV int warp_min_i(int v) {
#ifdef __CUDA_ARCH__
  return __reduce_min_sync(kFullMask, v);
#else
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int t = __shfl_xor_sync(kFullMask, v, o);
    v = t < v ? t : v;
  }
  return v;
#endif
}

// Q8 log2 of a non-zero magnitude (nsx_core.c:361-367)
NSB_DEV int nsx_log2_q8(unsigned v, const int16_t* log_frac) {
  const int zeros = fx_norm_u32(v);
  const int frac = (int)(((v << zeros) & 0x7FFFFFFFu) >> 23);
  return fx_s16(((31 - zeros) << 8) + log_frac[frac]);
}

// kIndicatorTable interpolation (nsx_core_c.c:104-113 and twins)
NSB_DEV int nsx_indicator(const int16_t* tab, unsigned x_q14, bool rounded) {
  const int idx = fx_s16((int)(x_q14 >> 14));
  const int base = tab[idx];
  const int step = fx_s16(tab[idx + 1] - tab[idx]);
  const int frac = (int)(x_q14 & 0x3fffu);
  return fx_s16(base + fx_s16(rounded ? fx_mul_rsft_round(step, frac, 14) : ((step * frac) >> 14)));
}

// 2^(num - exp*log2(bin)) in Q(minNorm-stages) (nsx_core.c:585-627); est/est_avg
// stay as passed when the exponent is not positive.
NSB_DEV void nsx_parametric(int exp_avg, int num_avg, int log_idx_bin, int min_norm, int stages,
                            int block_index, unsigned& est, unsigned& est_avg) {
  int t2 = (exp_avg * log_idx_bin) >> 15;
  int t1 = num_avg - t2;
  t1 += fx_shl(min_norm - stages, 11);
  if (t1 > 0) {
    const int int_part = fx_s16(t1 >> 11);
    const int frac = fx_s16(t1 & 0x7ff);
    if (frac >> 10) {
      t2 = (2048 - frac) * 1244;
      t2 = 2048 - (t2 >> 10);
    } else {
      t2 = (frac * 804) >> 10;
    }
    t2 = fx_shift_w32(t2, int_part - 11);
    est_avg = (unsigned)fx_shl(1, int_part) + (unsigned)t2;
    est = est_avg * (unsigned)(block_index + 1);
  }
}

// Two highest histogram peaks with the sequential-scan semantics of
// nsx_core.c:923-939 (strict '>' both times), warp parallel; positions 2i+1.
NSB_DEV void nsx_two_peaks(const int* h, int lane, int& w1, unsigned& p1, int& w2, unsigned& p2) {
  int bv = 0, bi = 0x7fffffff;
  for (int i = lane; i < 1000; i += 32) {
    const int v = __ldcg(h + i);
    if (v > bv) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int ov = __shfl_xor_sync(kFullMask, bv, o), oi = __shfl_xor_sync(kFullMask, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  const int i1 = bi;
  w1 = bv;
  p1 = bv > 0 ? (unsigned)(2 * i1 + 1) : 0u;
  int cv = 0, ci = 0x7fffffff;
  for (int i = lane; i < 1000; i += 32) {
    const int v = __ldcg(h + i);
    if (i != i1 && v > cv) { cv = v; ci = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int ov = __shfl_xor_sync(kFullMask, cv, o), oi = __shfl_xor_sync(kFullMask, ci, o);
    if (ov > cv || (ov == cv && oi < ci)) { cv = ov; ci = oi; }
  }
  w2 = cv;
  p2 = cv > 0 ? (unsigned)(2 * ci + 1) : 0u;
}

// Threshold / weight re-estimation every 512 frames (nsx_core.c:870-1014).
NSB_DEV void nsx_extract_params(int* Hw, int* hist, int stages, int max_lrt, int min_lrt, int lane) {
  int* hLrt = hist;
  int* hFlat = hist + 1000;
  int* hDiff = hist + 2000;
  int avg = 0, avg_sq = 0, avg_compl = 0, num = 0;
  for (int i = lane; i < 1000; i += 32) {
    const int c = __ldcg(hLrt + i);
    const int j = 2 * i + 1;
    const int t = c * j;
    if (i < 10) {
      avg += t;
      num += c;
    }
    avg_compl += t;
    avg_sq += t * j;
  }
  avg = warp_sum_i(avg);
  avg_sq = warp_sum_i(avg_sq);
  avg_compl = warp_sum_i(avg_compl);
  num = fx_s16(warp_sum_i(num));
  const int fluct = avg_sq * num - avg * avg_compl;
  const int thr_fluct = 10240 * num;
  const unsigned tu = 6u * (unsigned)avg;
  int use_diff = 1;
  if (fluct < thr_fluct || num == 0 || tu > (unsigned)(100 * num)) {
    Hw[kX_thrLrt] = max_lrt;
  } else {
    const int t = (int)((tu << (9 + stages)) / (unsigned)num / 25u);
    Hw[kX_thrLrt] = t > max_lrt ? max_lrt : (t < min_lrt ? min_lrt : t);
  }
  if (fluct < thr_fluct) use_diff = 0;

  int w1, w2;
  unsigned p1, p2;
  nsx_two_peaks(hFlat, lane, w1, p1, w2, p2);
  int use_flat = 1;
  if ((p1 - p2 < 4u) && (w2 * 2 > w1)) {
    w1 += w2;
    p1 = (p1 + p2) >> 1;
  }
  if (w1 < 154 || p1 < 24u) {
    use_flat = 0;
  } else {
    const unsigned v = 922u * p1;
    Hw[kX_thrFlat] = (int)(v > 38912u ? 38912u : (v < 4096u ? 4096u : v));
  }
  if (use_diff) {
    nsx_two_peaks(hDiff, lane, w1, p1, w2, p2);
    if ((p1 - p2 < 4u) && (w2 * 2 > w1)) {
      w1 += w2;
      p1 = (p1 + p2) >> 1;
    }
    const unsigned v = 6u * p1;
    Hw[kX_thrDiff] = (int)(v > 100u ? 100u : (v < 16u ? 16u : v));
    if (w1 < 154) use_diff = 0;
  }
  const int fsum = 6 / (1 + use_flat + use_diff);
  Hw[kX_wLrt] = fsum;
  Hw[kX_wFlat] = use_flat * fsum;
  Hw[kX_wDiff] = use_diff * fsum;
  __syncwarp();
  for (int i = lane; i < 3000; i += 32) hist[i] = 0;
}

template <int ANA, int NB>
__global__ void __launch_bounds__(kNsxMaxWarpsPerCta * 32, 1)
nsx_process_kernel(const NsxLaunch p) {
  constexpr int LANES = ANA / 8;
  constexpr int HALF = ANA / 2;
  constexpr int NBIN = HALF + 1;
  constexpr int STAGES = ANA == 256 ? 8 : 7;
  constexpr int NSLOT = 5;
  extern __shared__ uint4 nsx_smem4[];
  uint32_t* smem = reinterpret_cast<uint32_t*>(nsx_smem4);
  int16_t* s_win = reinterpret_cast<int16_t*>(smem);
  const int2* s_tw = reinterpret_cast<const int2*>(smem + kNsxImgFftTw);   // FFT twiddles, regrouped per stage
  int16_t* s_logf = reinterpret_cast<int16_t*>(smem + 256);

  const int lane = lane_id();
  const int warp = __shfl_sync(kFullMask, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform: see nsf_kernel.cuh
  const NsxTables* T = p.tables;
  const int warps_per_cta = (int)(blockDim.x >> 5);
  const int sidx = (int)blockIdx.x * warps_per_cta + warp;
  const bool live = sidx < p.n_streams;

  uint32_t* W = smem + kNsxCtaTableWords + warp * kNsxWarpWords;
  int* Hr = reinterpret_cast<int*>(W);                   // header, double buffered (see nsf_kernel.cuh)
  int* Hw = reinterpret_cast<int*>(W + kNsxHdrWords);
  uint4* RA = reinterpret_cast<uint4*>(W + 2 * kNsxHdrWords);
  uint4* RB = RA + 129;
  uint32_t* scr = reinterpret_cast<uint32_t*>(RB + 129);  // 512 words: FFT transposes
  uint32_t* buf = scr + 512;                              // 136 words: time / spectrum staging

  // ---- tables and state: HBM -> shared by TMA bulk copies issued by one lane (see nsf_kernel.cuh);
  // the sample histories pass through the FFT scratch on their way to registers
  mbar_t* bars = reinterpret_cast<mbar_t*>(smem + kNsxTableImgWords);
  mbar_t* barT = bars;
  mbar_t* barH = bars + 1 + 2 * warp;   // header + sample histories
  mbar_t* barB = barH + 1;              // per-bin records
  if (threadIdx.x == 0) mbar_init(barT, 1);
  if (lane == 0) {
    mbar_init(barH, 1);
    mbar_init(barB, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect(barT, sizeof(uint32_t) * kNsxTableImgWords);
    bulk_load(smem, T->img[ANA == 256 ? 0 : 1], sizeof(uint32_t) * kNsxTableImgWords, barT);
  }
  const int slot = __shfl_sync(kFullMask, live ? (p.slots ? p.slots[sidx] : p.slot_base + sidx) : 0, 0);
  uint32_t* gS = p.state + (size_t)slot * kNsxStateWords;
  int* gHist = reinterpret_cast<int*>(gS + kNsxOffHist);
  uint32_t* gInit = gS + kNsxOffInitMagn;
  const bool act = lane < LANES;
  constexpr unsigned kHistBytes = LANES * 3 * sizeof(int16_t);   // 192 / 96
  constexpr unsigned kRecBytes = NBIN * sizeof(uint4);           // 2064 / 1040
  if (live && lane == 0) {
    mbar_arrive_expect(barH, kNsxHdrWords * 4 + (2 + (NB - 1)) * kHistBytes);
    bulk_load(Hr, gS, kNsxHdrWords * 4, barH);
    bulk_load(scr, gS + kNsxOffAna, kHistBytes, barH);
    bulk_load(scr + 48, gS + kNsxOffSyn, kHistBytes, barH);
#pragma unroll
    for (int b = 0; b < NB - 1; ++b) bulk_load(scr + 96 + 48 * b, gS + kNsxOffHb + 48 * b, kHistBytes, barH);
    mbar_arrive_expect(barB, 2 * kRecBytes);
    bulk_load(RA, gS + kNsxOffRecA, kRecBytes, barB);
    bulk_load(RB, gS + kNsxOffRecB, kRecBytes, barB);
  }
  // prefetch across CTAs (see nsf_kernel.cuh): state and first frame of the stream that will run in
  // this warp's place about half a wave from now, into L2
  if (live && p.prefetch_ahead > 0 && sidx + p.prefetch_ahead < p.n_streams) {
    const int an = sidx + p.prefetch_ahead;
    const uint32_t* aS = p.state + (size_t)(p.slots ? p.slots[an] : p.slot_base + an) * kNsxStateWords;
    if (lane == 0) {
      bulk_prefetch_l2(aS, (kNsxOffHb + 48 * (NB - 1)) * 4);   // header | histories | high-band delay
      bulk_prefetch_l2(aS + kNsxOffRecA, kRecBytes);
      bulk_prefetch_l2(aS + kNsxOffRecB, kRecBytes);
    }
    if (p.frames > 0) {
      const char* a = reinterpret_cast<const char*>(static_cast<const int16_t*>(p.in) + (size_t)an * (size_t)p.in_stream_stride);
      const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(a) & 127u);
      if (lane * 128u < mis + (ANA == 256 ? 320u : 160u)) line_prefetch_l2(a - mis + 128 * lane);
    }
  }

  const int16_t* gin = static_cast<const int16_t*>(p.in) + (size_t)sidx * (size_t)p.in_stream_stride;
  int16_t* gout = static_cast<int16_t*>(p.out) + (size_t)sidx * (size_t)p.out_stream_stride;

  // frame samples lane + LANES*r, r < 5, per band; prefetched one frame ahead
  int cur[NB][5];
  auto load_frame = [&](int f, int (&dst)[NB][5]) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int16_t* src = gin + (size_t)f * (size_t)p.in_frame_stride + (size_t)b * (size_t)p.in_band_stride;
#pragma unroll
      for (int r = 0; r < 5; ++r) dst[b][r] = act ? (int)src[lane + LANES * r] : 0;
    }
  };
  // the first frame's PCM goes out with the state copies, before anything is waited for
  if (live && p.frames > 0) load_frame(0, cur);

  mbar_wait_cta(barT, 0);
  if (!live) return;
  mbar_wait_warp(barH, 0);
  // (waiting for the records here rather than before their first use costs nothing measurable --
  // both copies are in flight together -- and keeps a pointer and a flag out of the frame loop's
  // 72 registers)
  mbar_wait_warp(barB, 0);
  int ana_h[3], syn_h[3], hb_h[NB > 1 ? NB - 1 : 1][3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    ana_h[r] = syn_h[r] = 0;
    if (act) {
      ana_h[r] = reinterpret_cast<const int16_t*>(scr)[lane + LANES * r];
      syn_h[r] = reinterpret_cast<const int16_t*>(scr + 48)[lane + LANES * r];
    }
#pragma unroll
    for (int b = 0; b < NB - 1; ++b)
      hb_h[b][r] = act ? reinterpret_cast<const int16_t*>(scr + 96 + 48 * b)[lane + LANES * r] : 0;
  }
  __syncwarp();

  const int overdrive = Hr[kX_overdrive];
  const int denoise_bound = Hr[kX_denoiseBound];
  const int gain_map = Hr[kX_gainMap];
  const int mode = Hr[kX_mode];
  const int max_lrt = ANA == 256 ? 0x0080000 : 0x0040000;
  const int min_lrt = ANA == 256 ? 104858 : 52429;

  // optional lock step of the CTA's warps (full CTAs only): warps in the same phase share
  // instruction-cache lines
  const bool cta_sync = NSX_FRAME_SYNC && ((int)blockIdx.x + 1) * warps_per_cta <= p.n_streams;
#define NSX_PHASE_SYNC() do { if (NSX_FRAME_SYNC >= 2 && cta_sync) __syncthreads(); } while (0)
  for (int f = 0; f < p.frames; ++f) {
    if (cta_sync) __syncthreads();
    int nxt[NB][5] = {};
    if (f + 1 < p.frames) load_frame(f + 1, nxt);
    Hw[lane] = Hr[lane];

    // ---- AnalysisUpdate + window (nsx_core.c:523-540); x[r] = analysisBuffer[lane+LANES*r]
    int x[8], wd[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r] = r < 3 ? ana_h[r] : cur[0][r - 3];
#pragma unroll
    for (int r = 0; r < 3; ++r) ana_h[r] = x[r + 5];
    int max_int = 0, max_i16 = -1;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      // (window <= 16384 = Q14 of 1.0 and |x| <= 32768: the rounded product already is an int16, no cast needed)
      wd[r] = act ? fx_mul_rsft_round(s_win[act ? lane + LANES * r : 0], x[r], 14) : 0;
      const int a = wd[r] < 0 ? -wd[r] : wd[r];
      max_int = a > max_int ? a : max_int;
      const int a16 = wd[r] > 0 ? wd[r] : fx_s16(-wd[r]);  // get_scaling_square.c:33 (int16 negate)
      if (act) max_i16 = a16 > max_i16 ? a16 : max_i16;
    }
    max_int = warp_max_i(max_int);
    max_i16 = warp_max_i(max_i16);
    // WebRtcSpl_Energy (energy.c:20, get_scaling_square.c:20)
    int scale_in;
    {
      const int nbits = ANA == 256 ? 9 : 8;  // GetSizeInBits(anaLen)
      if (max_i16 == 0) {
        scale_in = 0;
      } else {
        const int t = fx_norm_w32(max_i16 * max_i16);
        scale_in = t > nbits ? 0 : nbits - t;
      }
    }
    int energy_in = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) energy_in += (wd[r] * wd[r]) >> scale_in;
    energy_in = warp_sum_i(energy_in);
    const int max_win = max_int > 32767 ? 32767 : max_int;
    const int norm_data = fx_norm_w16(max_win);
    const bool zero_input = (max_win == 0);

    int outv[5];           // band-0 output samples lane + LANES*r
    int hb_gain = 16384;
    bool hb_apply = false;

    if (zero_input) {
      // nsx_core.c:1439-1452: flush the overlap, no statistics update
#pragma unroll
      for (int r = 0; r < 5; ++r) outv[r] = r < 3 ? syn_h[r] : 0;
#pragma unroll
      for (int r = 0; r < 3; ++r) syn_h[r] = 0;
      // as many barriers as the other branch (warps of one CTA may take different branches)
      NSX_PHASE_SYNC(); NSX_PHASE_SYNC(); NSX_PHASE_SYNC(); NSX_PHASE_SYNC(); NSX_PHASE_SYNC();
    } else {
      const int block_index = Hr[kX_blockIndex] + 1;
      Hw[kX_blockIndex] = block_index;
      int min_norm = Hr[kX_minNorm];
      const int net_norm = STAGES - norm_data;
      int rs_magn = norm_data - min_norm;
      const int rs_init = -rs_magn > 0 ? -rs_magn : 0;
      min_norm -= rs_init;
      Hw[kX_minNorm] = min_norm;
      if (rs_magn < 0) rs_magn = 0;
      const int q_magn = norm_data - STAGES;

      // ---- forward FFT: NormalizeRealBuffer (:543) -> bit reversal -> radix-2 stages
      int vr[8], vi[8];
      {
        // position 8*lane + r holds sample bitrev(8*lane + r) = rev5(lane) + LANES*rev3(r):
        // the samples of lane rev(lane), in order rev3(r)
        const int src_lane = ANA == 256 ? (int)(__brev((unsigned)lane) >> 27) : (int)(__brev((unsigned)lane) >> 28);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int rr = ((r & 1) << 2) | (r & 2) | ((r >> 2) & 1);
          const int nv = fx_s16(fx_shl(wd[rr], norm_data));
          vr[r] = __shfl_sync(kFullMask, nv, src_lane & 31);
          vi[r] = 0;
        }
      }
      fx_warp_cfft<false, ANA>(vr, vi, scr, s_tw, lane);

      // ---- spectrum, magnitude (nsx_core.c:1247-1272)
      int re[NSLOT], im[NSLOT];
      unsigned magn[NSLOT];
      const int nyq_re = __shfl_sync(kFullMask, vr[4], 0);  // position HALF sits in lane 0, r = 4
      unsigned magn_energy = 0, sum_magn = 0;
      bool any_zero = false;
#pragma unroll
      for (int j = 0; j < NSLOT; ++j) {
        const bool nyq = j == NSLOT - 1;
        const int wre = nyq ? nyq_re : vr[j], wim = nyq ? 0 : vi[j];
        const int k = nyq ? HALF : lane + LANES * j;
        re[j] = wre;
        im[j] = fx_s16(-wim);
        unsigned e;
        if (nyq || k == 0) {
          im[j] = 0;
          e = (unsigned)(re[j] * re[j]);
          magn[j] = (unsigned)(re[j] < 0 ? -re[j] : re[j]) & 0xffffu;
        } else {
          e = (unsigned)(wre * wre) + (unsigned)(wim * wim);
          magn[j] = fx_sqrt_floor(e) & 0xffffu;
        }
        const bool mine = nyq ? (lane == 0) : act;
        if (mine) {
          magn_energy += e;
          sum_magn += magn[j];
          if (k >= 1 && magn[j] == 0u) any_zero = true;
        }
      }
      magn_energy = warp_sum_u(magn_energy);
      sum_magn = warp_sum_u(sum_magn);

      // ---- start-up statistics (nsx_core.c:1273-1418)
      int pink_exp = Hr[kX_pinkExp], pink_num = Hr[kX_pinkNum];
      unsigned white = (unsigned)Hr[kX_whiteLevel];
      unsigned init_m[NSLOT];
#pragma unroll
      for (int j = 0; j < NSLOT; ++j) init_m[j] = 0u;
      if (block_index < 50) {
        int slm = 0, slilm = 0;
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const bool nyq = j == NSLOT - 1;
          const int k = nyq ? HALF : lane + LANES * j;
          const bool mine = nyq ? (lane == 0) : act;
          unsigned im0 = (mine || nyq) ? gInit[k] : 0u;
          im0 >>= rs_init;
          im0 += magn[j] >> rs_magn;
          init_m[j] = im0;
          if (mine) {
            gInit[k] = im0;
            if (k >= 5) {
              const int l2 = magn[j] ? nsx_log2_q8(magn[j], s_logf) : 0;
              slm += l2;
              slilm += (T->log_idx[k] * l2) >> 3;
            }
          }
        }
        slm = warp_sum_i(slm);
        slilm = warp_sum_i(slilm);
        white >>= rs_init;
        {
          unsigned tu = sum_magn * (unsigned)overdrive;
          tu >>= STAGES + 8;
          tu >>= rs_magn;
          white += tu;
        }
        Hw[kX_whiteLevel] = (int)white;
        int det = T->det5, sum_log_i = T->sum_log_idx5, sum_log_i_sq = T->sum_sq_log_idx5;
        if (ANA == 128) {
          int t1 = det;
          t1 += (T->sum_log_idx65 * sum_log_i) >> 9;
          t1 -= (T->sum_log_idx65 * T->sum_log_idx65) >> 10;
          t1 -= fx_shl(sum_log_i_sq, 4);
          t1 -= (fx_s16(NBIN - 5) * T->sum_sq_log_idx65) >> 2;
          det = fx_s16(t1);
          sum_log_i = fx_s16(sum_log_i - T->sum_log_idx65);
          sum_log_i_sq = fx_s16(sum_log_i_sq - T->sum_sq_log_idx65);
        }
        int zeros = 16 - fx_norm_w32(slm);
        if (zeros < 0) zeros = 0;
        int t1 = fx_shl(slm, 1);
        const unsigned slm_u16 = (unsigned)(t1 >> zeros) & 0xffffu;
        int t2 = sum_log_i_sq * (int)slm_u16;
        unsigned tu = (unsigned)(slilm >> 12);
        unsigned tu16 = ((unsigned)sum_log_i << 1) & 0xffffu;
        if ((unsigned)sum_log_i > tu) tu16 >>= zeros;
        else tu >>= zeros;
        t2 -= (int)(tu * tu16);
        det >>= zeros;
        t2 = det != 0 ? t2 / det : 0x7FFFFFFF;
        t2 += fx_shl(net_norm, 11);
        if (t2 < 0) t2 = 0;
        pink_num += t2;
        Hw[kX_pinkNum] = pink_num;
        t2 = sum_log_i * (int)slm_u16;
        t1 = slilm >> (3 + zeros);
        t1 *= NBIN - 5;
        t2 -= t1;
        if (t2 > 0) {
          t1 = det != 0 ? t2 / det : 0x7FFFFFFF;
          pink_exp += t1 > 16384 ? 16384 : (t1 < 0 ? 0 : t1);
        }
        Hw[kX_pinkExp] = pink_exp;
      }

      // ---- ComputeSpectralFlatness (nsx_core.c:1021-1083)
      unsigned feat_flat = (unsigned)Hr[kX_featFlat];
      {
        unsigned num = 0;
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const bool nyq = j == NSLOT - 1;
          const int k = nyq ? HALF : lane + LANES * j;
          const bool mine = nyq ? (lane == 0) : act;
          if (mine && k >= 1 && magn[j]) num += (unsigned)nsx_log2_q8(magn[j], s_logf);
        }
        num = warp_sum_u(num);
        const bool zero_bin = __ballot_sync(kFullMask, any_zero) != 0u;
        if (zero_bin) {
          feat_flat -= (feat_flat * 4915u) >> 14;
        } else {
          const unsigned den = sum_magn - __shfl_sync(kFullMask, magn[0], 0);
          const int zeros = fx_norm_u32(den);
          const int frac = (int)(((den << zeros) & 0x7FFFFFFFu) >> 23);
          int t = ((31 - zeros) << 8) + s_logf[frac];
          int lcur = (int)num;
          lcur += fx_shl(STAGES - 1, STAGES + 7);
          lcur -= fx_shl(t, STAGES - 1);
          lcur = fx_shl(lcur, 10 - STAGES);
          t = (int)(0x00020000 | ((lcur >= 0 ? lcur : -lcur) & 0x0001FFFF));
          const int int_part = 7 - (lcur >> 17);
          const int curf = int_part > 0 ? (t >> int_part) : fx_shl(t, -int_part);
          t = curf - (int)feat_flat;
          t *= 4915;
          feat_flat += (unsigned)(t >> 14);
        }
        Hw[kX_featFlat] = (int)feat_flat;
      }

      NSX_PHASE_SYNC();
      // ---- NoiseEstimation (nsx_core.c:334-452)
      unsigned noise[NSLOT];
      int q_noise = Hr[kX_qNoise];
      {
        const int tabind = STAGES - norm_data;
        const int logval = tabind < 0 ? -T->log_tab[-tabind] : T->log_tab[tabind];
        int cnt[3], cdiv[3], cprod[3];
        bool latch[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          cnt[s] = Hr[kX_counter + s];
          cdiv[s] = T->counter_div[cnt[s]];
          cprod[s] = fx_s16(cnt[s] * cdiv[s]);
          latch[s] = cnt[s] >= 200;
          Hw[kX_counter + s] = (latch[s] ? 0 : cnt[s]) + 1;
        }
        // loop-invariant over the bins: the density increment and the start-up step
        int bb[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) bb[s] = fx_s16(fx_mul_rsft_round(21845, cdiv[s], 15));
        const int dsmall = block_index < 200 ? 1024 : 5120;
        // which estimator (if any) is latched into noiseEstQuantile this frame
        int sel = -1;
        if (block_index >= 200) {
#pragma unroll
          for (int s = 0; s < 3; ++s)
            if (latch[s]) sel = s;
        } else {
          sel = 2;
        }
        int lqs[NSLOT], quant[NSLOT], mx = -32768;
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const bool nyq = j == NSLOT - 1;
          const int k = nyq ? HALF : lane + LANES * j;
          const bool mine = nyq ? (lane == 0) : act;
          const uint4 a = RA[(mine || nyq) ? k : 0];
          int lq[3] = {fx_lo(a.x), fx_hi(a.x), fx_lo(a.y)};
          int dn[3] = {fx_hi(a.y), fx_lo(a.z), fx_hi(a.z)};
          quant[j] = fx_lo(a.w);
          // Branch-free: every data-dependent `if` of the reference's loop body (nsx_core.c:361-430) is a select
          // here, so that the 15 tracker updates of a lane form one block the scheduler can interleave.
          // fx_norm_u32(0) = 0 and the table entry of a zero fraction is 0, so the logarithm of magn = 0 is
          // computed harmlessly and discarded.
          const int l2 = nsx_log2_q8(magn[j], s_logf);
          const int lmagn = magn[j] ? fx_s16(fx_s16((l2 * 22713) >> 15) + logval) : logval;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            // density > 512: delta = FACTOR_Q7 >> (14 - norm16(density)) = 2621440 >> (31 - clz(density))
            const int dv = dn[s] > 513 ? dn[s] : 513;
            const int delta = dn[s] > 512 ? 2621440 >> (31 - __clz(dv)) : dsmall;   // shift >= 9: at most 5120
            // delta in [160, 5120] and counter_div in (0, 32767]: 0 <= t16 <= 10240, so the reference's
            // (int16_t) casts of t16 + 1, t16 + 2 and their halves / quarters (exact shifts) are identities
            const int t16 = (delta * cdiv[s]) >> 14;
            const int lq_up = fx_s16(lq[s] + ((t16 + 2) >> 2));
            const int lq_dn0 = fx_s16(lq[s] - ((((t16 + 1) >> 1) * 3) >> 1));
            const int lq_dn = lq_dn0 < logval ? logval : lq_dn0;
            lq[s] = lmagn > lq[s] ? lq_up : lq_dn;
            const int d = fx_s16(lmagn - lq[s]);
            const int aa = fx_mul_rsft_round(dn[s], cprod[s], 15);   // two int16 factors: within [-32767, 32766]
            dn[s] = (d >= 0 ? d : -d) < 3 ? fx_s16(aa + bb[s]) : dn[s];
          }
          lqs[j] = sel == 0 ? lq[0] : (sel == 1 ? lq[1] : lq[2]);
          if (mine && sel >= 0 && lqs[j] > mx) mx = lqs[j];
          if (mine) {
            uint4 o;
            o.x = fx_pack(lq[0], lq[1]);
            o.y = fx_pack(lq[2], dn[0]);
            o.z = fx_pack(dn[1], dn[2]);
            o.w = a.w;
            RA[k] = o;
          }
        }
        if (sel >= 0) {
          // UpdateNoiseEstimate (:303-331): Q-domain from the largest log quantile
          mx = warp_max_i(mx);
          q_noise = 14 - fx_mul_rsft_round(11819, mx, 21);
          Hw[kX_qNoise] = q_noise;
#pragma unroll
          for (int j = 0; j < NSLOT; ++j) {
            const int e = 11819 * lqs[j];
            int m = 0x00200000 | (e & 0x001FFFFF);
            int sh = fx_s16(e >> 21);
            sh = fx_s16(sh - 21);
            sh = fx_s16(sh + q_noise);
            if (sh < 0) m >>= -sh;
            else m = fx_shl(m, sh);
            quant[j] = fx_sat16(m);
          }
        }
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) noise[j] = (unsigned)quant[j];
        // quantile is written back together with the filter below
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) lqs[j] = quant[j];
        // keep quant in lqs for the record update
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const bool nyq = j == NSLOT - 1;
          const int k = nyq ? HALF : lane + LANES * j;
          const bool mine = nyq ? (lane == 0) : act;
          if (mine) reinterpret_cast<int16_t*>(&RA[k])[6] = (int16_t)quant[j];
        }
      }

      // ---- previous noise, parametric start-up model (nsx_core.c:1611-1711)
      uint4 rb[NSLOT];
      unsigned filter_prev[NSLOT], filter_tmp[NSLOT];
#pragma unroll
      for (int j = 0; j < NSLOT; ++j) {
        const bool nyq = j == NSLOT - 1;
        const int k = nyq ? HALF : lane + LANES * j;
        const bool mine = nyq ? (lane == 0) : act;
        rb[j] = RB[(mine || nyq) ? k : 0];
        filter_prev[j] = (unsigned)reinterpret_cast<const uint16_t*>(&RA[(mine || nyq) ? k : 0])[7];
        filter_tmp[j] = (unsigned)denoise_bound;
      }
      if (block_index < 50) {
        const int q_use = q_noise < min_norm - STAGES ? q_noise : min_norm - STAGES;
        int exp_avg = 0, num_avg = 0;
        unsigned est0 = 0, est0_avg = 0;
        if (pink_exp) {
          const int dv = fx_s16(block_index + 1);
          exp_avg = fx_s16(dv != 0 ? pink_exp / dv : 0x7FFFFFFF);
          num_avg = dv != 0 ? pink_num / dv : 0x7FFFFFFF;
          nsx_parametric(exp_avg, num_avg, T->log_idx[5], min_norm, STAGES, block_index, est0, est0_avg);
        } else {
          est0 = white;
          est0_avg = est0 / (unsigned)(block_index + 1);
        }
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const bool nyq = j == NSLOT - 1;
          const int k = nyq ? HALF : lane + LANES * j;
          unsigned est = est0, est_avg = est0_avg;
          if (pink_exp && k >= 5) {
            est = 0;
            est_avg = 0;
            nsx_parametric(exp_avg, num_avg, T->log_idx[k < NBIN ? k : 0], min_norm, STAGES, block_index, est, est_avg);
          }
          if (init_m[j]) {
            unsigned u1 = est * (unsigned)overdrive;
            unsigned numer = init_m[j] << 8;
            if (numer > u1) {
              numer -= u1;
              int ns = fx_norm_u32(numer);
              ns = ns > 6 ? 6 : (ns < 0 ? 0 : ns);
              numer <<= ns;
              u1 = init_m[j] >> (6 - ns);
              if (u1 == 0) u1 = 1;
              const unsigned u2 = numer / u1;
              filter_tmp[j] = u2 > 16384u ? 16384u : (u2 < (unsigned)denoise_bound ? (unsigned)denoise_bound : u2);
              filter_tmp[j] &= 0xffffu;
            }
          }
          unsigned u1 = noise[j] >> (q_noise - q_use);
          unsigned u2 = est_avg >> (min_norm - STAGES - q_use);
          int nsh = 0;
          if (u1 & 0xfc000000u) {
            u1 >>= 6;
            u2 >>= 6;
            nsh = 6;
          }
          u1 *= (unsigned)block_index;
          u2 *= (unsigned)(50 - block_index);
          noise[j] = (u1 + u2) / 50u;
          noise[j] <<= nsh;
        }
        q_noise = q_use;
      }
      unsigned time_avg = (unsigned)Hr[kX_timeAvgEnergy];
      if (block_index < 200) {
        const unsigned tmp = (unsigned)Hr[kX_timeAvgEnergyTmp] + (magn_energy >> (2 * norm_data + STAGES - 1));
        Hw[kX_timeAvgEnergyTmp] = (int)tmp;
        const unsigned dv = (unsigned)(block_index + 1) & 0xffffu;
        time_avg = dv ? tmp / dv : 0xFFFFFFFFu;
        Hw[kX_timeAvgEnergy] = (int)time_avg;
      }

      NSX_PHASE_SYNC();
      // ---- step 1: post / prior SNR (nsx_core.c:1723-1786)
      const unsigned sat_max = 1048575u;
      const int prev_q_magn = Hr[kX_prevQMagn], prev_q_noise = Hr[kX_prevQNoise];
      unsigned post_snr[NSLOT], prior_snr[NSLOT], prev_near[NSLOT];
      {
        const int post_shifts = 6 + q_magn - q_noise;
        const int nsh = 5 - prev_q_magn + prev_q_noise;
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          // (branch-free: the divisions run on a guarded divisor and the reference's cases select the result,
          // so that the five slots' division chains interleave)
          unsigned u1 = magn[j] << 6;
          unsigned u2 = post_shifts < 0 ? (noise[j] >> -post_shifts) : (noise[j] << post_shifts);
          {
            const unsigned q = fx_udiv_q20(u1 << 11, u2 ? u2 : 1u);   // capped at sat_max below
            const unsigned capped = (u2 && q < sat_max) ? q : sat_max;
            post_snr[j] = u1 > u2 ? capped : 2048u;
          }
          const unsigned prev_magn = rb[j].w & 0xffffu;
          const unsigned near_est = prev_magn * filter_prev[j];
          u1 = near_est << 3;
          u2 = rb[j].z >> nsh;
          {
            const unsigned q = fx_udiv_q20(u1, u2 ? u2 : 1u);   // capped at sat_max below
            prev_near[j] = (u2 && q < sat_max) ? q : sat_max;
          }
          const unsigned prior = prev_near[j] * 2007u + (post_snr[j] - 2048u) * 41u + 512u;
          prior_snr[j] = 2048u + (prior >> 10);
        }
      }

      // ---- ComputeSpectralDifference (nsx_core.c:1090-1180)
      unsigned feat_diff = (unsigned)Hr[kX_featDiff];
      unsigned cur_avg_energy = (unsigned)Hr[kX_curAvgEnergy];
      {
        int sum_p = 0, max_p = 0, min_p = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const bool mine = (j == NSLOT - 1) ? (lane == 0) : act;
          if (mine) {
            const int pv = (int)rb[j].y;
            sum_p += pv;
            max_p = pv > max_p ? pv : max_p;
            min_p = pv < min_p ? pv : min_p;
          }
        }
        sum_p = warp_sum_i(sum_p);
        max_p = warp_max_i(max_p);
        min_p = warp_min_i(min_p);
        const int avg_pause = sum_p >> (STAGES - 1);
        const int avg_magn = (int)(sum_magn >> (STAGES - 1));
        int t1 = (max_p - avg_pause) > (avg_pause - min_p) ? (max_p - avg_pause) : (avg_pause - min_p);
        int n_shifts = 10 + STAGES - fx_norm_w32(t1);
        if (n_shifts < 0) n_shifts = 0;
        unsigned var_m = 0, var_p = 0, cov_u = 0;
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const bool mine = (j == NSLOT - 1) ? (lane == 0) : act;
          if (mine) {
            const int d16 = fx_s16((int)magn[j] - avg_magn);
            const int t2 = (int)rb[j].y - avg_pause;
            var_m += (unsigned)(d16 * d16);
            cov_u += (unsigned)t2 * (unsigned)d16;
            const int t3 = t2 >> n_shifts;
            var_p += (unsigned)t3 * (unsigned)t3;
          }
        }
        var_m = warp_sum_u(var_m);
        var_p = warp_sum_u(var_p);
        const int cov = (int)warp_sum_u(cov_u);
        cur_avg_energy += magn_energy >> (2 * norm_data + STAGES - 1);
        unsigned diff = var_m;
        if (var_p && cov) {
          unsigned u1 = (unsigned)(cov >= 0 ? cov : -cov);
          const int norm32 = fx_norm_u32(u1) - 16;
          if (norm32 > 0) u1 <<= norm32;
          else u1 >>= -norm32;
          const unsigned u2 = u1 * u1;
          n_shifts += norm32;
          n_shifts <<= 1;
          if (n_shifts < 0) {
            var_p >>= -n_shifts;
            n_shifts = 0;
          }
          if (var_p > 0u) {
            u1 = u2 / var_p;
            u1 >>= n_shifts;
            diff -= diff < u1 ? diff : u1;
          } else {
            diff = 0;
          }
        }
        const unsigned u1 = diff >> (2 * norm_data);
        if (feat_diff > u1) feat_diff -= ((feat_diff - u1) * 77u) >> 8;
        else feat_diff += ((u1 - feat_diff) * 77u) >> 8;
      }

      // ---- histograms / threshold extraction (nsx_core.c:1796-1836, 820-1015)
      {
        const int cnt_upd = Hr[kX_cntThresUpdate] + 1;
        const bool flag = cnt_upd == 512;
        if (!flag) {
          if (lane == 0) {
            unsigned idx = (unsigned)Hr[kX_featLrt];
            if (idx < 1000u) atomicAdd(gHist + idx, 1);
            idx = (feat_flat * 5u) >> 8;
            if (idx < 1000u) atomicAdd(gHist + 1000 + idx, 1);
            idx = 1000u;
            if (time_avg > 0u) idx = ((feat_diff * 5u) >> STAGES) / time_avg;
            if (idx < 1000u) atomicAdd(gHist + 2000 + idx, 1);
          }
          Hw[kX_cntThresUpdate] = cnt_upd;
        } else {
          __threadfence_block();
          __syncwarp();
          nsx_extract_params(Hw, gHist, STAGES, max_lrt, min_lrt, lane);
          Hw[kX_cntThresUpdate] = 0;
          cur_avg_energy >>= 9;
          const unsigned u1 = (cur_avg_energy + time_avg + 1u) >> 1;
          if (u1 != time_avg && feat_diff && time_avg > 0u) {
            int norm1 = 0;
            unsigned u3 = u1;
            while (0xFFFF0000u & u3) { u3 >>= 1; norm1++; }
            unsigned u2 = feat_diff;
            while (0xFFFF0000u & u2) { u2 >>= 1; norm1++; }
            u3 = u3 * u2;
            u3 /= time_avg;
            if (fx_norm_u32(u3) < norm1) {
              feat_diff = 0x007FFFFFu;
            } else {
              const unsigned vv = u3 << norm1;
              feat_diff = vv < 0x007FFFFFu ? vv : 0x007FFFFFu;
            }
          }
          time_avg = u1;
          Hw[kX_timeAvgEnergy] = (int)time_avg;
          cur_avg_energy = 0;
        }
        Hw[kX_featDiff] = (int)feat_diff;
        Hw[kX_curAvgEnergy] = (int)cur_avg_energy;
        __syncwarp();
      }

      NSX_PHASE_SYNC();
      // ---- WebRtcNsx_SpeechNoiseProb (nsx_core_c.c:26-261)
      unsigned nonspeech[NSLOT];
      int prior_ns;
      {
        int ksum = 0;
        int lrt[NSLOT];
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const bool mine = (j == NSLOT - 1) ? (lane == 0) : act;
          int bessel = (int)post_snr[j];
          const int nt = fx_norm_u32(post_snr[j]);
          const unsigned num = post_snr[j] << nt;
          const unsigned den = nt > 10 ? (prior_snr[j] << (nt - 11)) : (prior_snr[j] >> (11 - nt));
          // post_snr is in [2048, 2^20), so nt >= 12, den = prior_snr << (nt - 11) >= 2^nt and the quotient < 2^20
          bessel = den ? bessel - (int)fx_udiv_q20(num, den ? den : 1u) : 0;
          const int zeros = fx_norm_u32(prior_snr[j]);
          int frac32 = (int)(((prior_snr[j] << zeros) & 0x7FFFFFFFu) >> 19);
          int t = (frac32 * frac32 * -43) >> 19;
          t += (frac32 * 5412) >> 12;   // frac32 < 2^12
          frac32 = t + 37;
          t = (((31 - zeros) << 12) + frac32) - (11 << 12);
          const int log_t = (t * 178) >> 8;
          const int half_sum = (log_t + (int)rb[j].x) / 2;
          lrt[j] = (int)rb[j].x + (bessel - half_sum);
          if (mine) ksum += lrt[j];
        }
        ksum = warp_sum_i(ksum);
        Hw[kX_featLrt] = (ksum * 10) >> (STAGES + 11);
        // priors may have been re-estimated a few lines up (synced): read Hw
        const int thr_lrt = Hw[kX_thrLrt], w_lrt = Hw[kX_wLrt], w_flat = Hw[kX_wFlat], w_diff = Hw[kX_wDiff];
        const unsigned thr_flat = (unsigned)Hw[kX_thrFlat], thr_diff = (unsigned)Hw[kX_thrDiff];
        int ind = 16384;
        int t1 = ksum - thr_lrt;
        int n_shifts = 7 - STAGES;
        if (t1 < 0) {
          ind = 0;
          t1 = -t1;
          n_shifts++;
        }
        t1 = fx_shift_w32(t1, n_shifts);
        {
          const int idx = fx_s16(t1 >> 14);
          if (idx < 16 && idx >= 0) {
            const int vv = nsx_indicator(T->indicator, (unsigned)t1, false);
            ind = fx_s16(ind == 0 ? 8192 - vv : 8192 + vv);
          }
        }
        int ind_prior = w_lrt * ind;
        if (w_flat) {
          unsigned u1 = feat_flat * 400u, u2;
          ind = 16384;
          u2 = thr_flat - u1;
          n_shifts = 4;
          if (thr_flat < u1) {
            ind = 0;
            u2 = u1 - thr_flat;
            n_shifts++;
          }
          u1 = (u2 << n_shifts) / 25u;
          if (fx_s16((int)(u1 >> 14)) < 16) {
            const int vv = nsx_indicator(T->indicator, u1, false);
            ind = fx_s16(ind ? 8192 + vv : 8192 - vv);
          }
          ind_prior += w_flat * ind;
        }
        if (w_diff) {
          unsigned u1 = 0, u2, u3;
          if (feat_diff) {
            int nt = fx_norm_u32(feat_diff);
            if (20 - STAGES < nt) nt = 20 - STAGES;
            u1 = feat_diff << nt;
            u2 = time_avg >> (20 - STAGES - nt);
            if (u2 > 0u) u1 /= u2;
            else u1 = 0x7fffffffu;
          }
          u3 = (thr_diff << 17) / 25u;
          u2 = u1 - u3;
          n_shifts = 1;
          ind = 16384;
          if (u2 & 0x80000000u) {
            ind = 0;
            u2 = u3 - u1;
            n_shifts--;
          }
          u1 = u2 >> n_shifts;
          if (fx_s16((int)(u1 >> 14)) < 16) {
            const int vv = nsx_indicator(T->indicator, u1, true);
            ind = fx_s16(ind ? 8192 + vv : 8192 - vv);
          }
          ind_prior += w_diff * ind;
        }
        const int ind16 = fx_s16((98307 - ind_prior) / 6);
        prior_ns = Hr[kX_priorNonSpeech];
        const int d16 = fx_s16(ind16 - prior_ns);
        prior_ns = fx_s16(prior_ns + fx_s16((1638 * d16) >> 14));
        Hw[kX_priorNonSpeech] = prior_ns;
        const int n2_prior = fx_norm_w16(16384 - prior_ns);
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          // nsx_core_c.c:214-258, branch-free.  The reference evaluates this only for priorNonSpeechProb > 0 and
          // logLrt < 65300; the clamp keeps the shifts of the discarded case in range.  n = n1 + n2 picks one
          // of the reference's two scalings, which differ only in their shift counts:
          //   n < 15:  inv >>= 15 - n;  inv = (inv * (16384 - prior)) >> (n - 7)
          //   n >= 15:                  inv = (inv * (16384 - prior)) >> 8
          // Both operands of the last division are positive (0 <= prior <= 16384 by construction of the
          // indicator mean, inv >= 0), so it is an unsigned division.
          const int lr = lrt[j] < 65299 ? lrt[j] : 65299;
          const int e = (lr * 23637) >> 14;
          int int_part = fx_s16(e >> 12);
          if (int_part < -8) int_part = -8;
          const int frac = e & 0xfff;
          int t2 = (frac * frac * 44) >> 19;
          t2 += (frac * 84) >> 7;
          int inv = fx_shl(1, 8 + int_part) + fx_shift_w32(t2, int_part - 4);
          const int nn = fx_norm_w32(inv) + n2_prior;
          const int pre = 15 - nn > 0 ? 15 - nn : 0;
          const int post = nn - 7 < 0 ? 0 : (nn - 7 > 8 ? 8 : nn - 7);
          inv = ((inv >> pre) * (16384 - prior_ns)) >> post;
          const bool live = prior_ns > 0 && lrt[j] < 65300 && nn >= 7;
          const unsigned den = live ? (unsigned)(prior_ns + inv) : 1u;
          const unsigned q = fx_udiv_q20((unsigned)fx_shl(prior_ns, 8), den);   // den >= prior_ns: quotient <= 256
          nonspeech[j] = live ? (q & 0xffffu) : 0u;
          rb[j].x = (uint32_t)lrt[j];
        }
      }

      // ---- step 2: noise update with gamma carried from bin k-1 (nsx_core.c:1841-1946)
      unsigned max_noise = 0;
      {
        const int post_shifts = prev_q_noise - q_magn;
        const int nsh = prev_q_magn - q_magn;
        const int nsh_up = nsh < 0 ? 0 : nsh;
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const bool nyq = j == NSLOT - 1;
          const bool mine = nyq ? (lane == 0) : act;
          // nonSpeechProb of bin k-1 decides the gamma this bin starts with
          unsigned pprev;
          if (nyq) {
            pprev = __shfl_sync(kFullMask, nonspeech[NSLOT - 2], LANES - 1);
          } else {
            pprev = __shfl_up_sync(kFullMask, nonspeech[j], 1);
            const unsigned wrap = __shfl_sync(kFullMask, nonspeech[j > 0 ? j - 1 : 0], LANES - 1);
            if (lane == 0) pprev = wrap;
          }
          const bool first = (!nyq && j == 0 && lane == 0);
          const unsigned gamma_in = (!first && pprev < 205u) ? 3u : 26u;
          const unsigned prev_noise16 = (rb[j].z >> 11) & 0xffffu;
          unsigned u2 = post_shifts < 0 ? (magn[j] >> -post_shifts) : (magn[j] << post_shifts);
          unsigned u1;
          int sign;
          if (prev_noise16 > u2) { sign = -1; u1 = prev_noise16 - u2; }
          else { sign = 1; u1 = u2 - prev_noise16; }
          // Branch-free (nsx_core.c:1868-1931).  With |delta| = 0 or nonSpeechProb = 0 the product u3 is 0 and the
          // update adds nothing, which is what the reference's `if` skips; when the bin's own gamma equals the
          // incoming one the second estimate IS the first, so the reference's "take the smaller" is a plain
          // minimum.
          const unsigned u3 = u1 * nonspeech[j];
          const bool big = (0x7c000000u & u3) != 0u;
          u2 = big ? (u3 >> 5) * gamma_in : (u3 * gamma_in) >> 5;
          unsigned upd = sign > 0 ? rb[j].z + u2 : rb[j].z - u2;
          const unsigned gamma = nonspeech[j] < 205u ? 3u : 26u;
          u2 = big ? (u3 >> 5) * gamma : (u3 * gamma) >> 5;
          u1 = sign > 0 ? rb[j].z + u2 : rb[j].z - u2;
          if (upd > u1) upd = u1;
          noise[j] = upd;
          if (mine && upd > max_noise) max_noise = upd;
          // magnitude during pauses (:1933-1945): both Q-domain cases are one expression in shifts that are
          // uniform over the frame
          const int t2 = fx_shift_w32((int)rb[j].y, -nsh);
          int t1 = fx_shl((int)magn[j], nsh_up) - (nsh < 0 ? t2 : (int)rb[j].y);
          t1 *= 13;
          t1 = (t1 + fx_shl(128, nsh_up)) >> (8 + nsh_up);
          rb[j].y = (uint32_t)(nonspeech[j] > 205u ? t2 + t1 : t2);
        }
      }
      max_noise = warp_max_u(max_noise);
      const int norm1 = fx_norm_u32(max_noise);
      const int q_noise_new = prev_q_noise + norm1 - 5;

      NSX_PHASE_SYNC();
      // ---- step 3: Wiener filter (nsx_core.c:1949-2015), state save (:2019-2031)
      unsigned hb_psum = 0, hb_gsum = 0;
      {
        const int nsh = prev_q_noise + 11 - q_magn;
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const bool nyq = j == NSLOT - 1;
          const int k = nyq ? HALF : lane + LANES * j;
          const bool mine = nyq ? (lane == 0) : act;
          unsigned cur_snr = 0, tm, tn;
          if (nsh < 0) {
            tm = magn[j];
            tn = noise[j] << -nsh;
          } else if (nsh > 17) {
            tm = magn[j] << 17;
            tn = noise[j] >> (nsh - 17);
          } else {
            tm = magn[j] << nsh;
            tn = noise[j];
          }
          {
            // (branch-free: computed on the wrapped difference and discarded unless tm > tn)
            unsigned u1 = tm - tn;
            int nn = fx_norm_u32(u1);
            if (nn > 11) nn = 11;
            u1 <<= nn;
            const unsigned u2 = tn >> (11 - nn);
            const unsigned q = u2 ? fx_udiv_q20(u1, u2 ? u2 : 1u) : u1;   // capped at sat_max below
            cur_snr = tm > tn ? (q < sat_max ? q : sat_max) : 0u;
          }
          const unsigned prior = prev_near[j] * 2007u + cur_snr * 41u;
          const unsigned dn = (unsigned)overdrive + ((prior + 8192u) >> 14);
          // dn = overdrive (>= 256) + ((prior + 8192) >> 14): the quotient stays below 2^14 + 1
          const unsigned f16 = fx_udiv_q20(prior + dn / 2u, dn) & 0xffffu;
          unsigned flt = f16 > 16384u ? 16384u : (f16 < (unsigned)denoise_bound ? (unsigned)denoise_bound : f16);
          if (block_index < 50) {
            const unsigned u1 = flt * (unsigned)block_index + filter_tmp[j] * (unsigned)(50 - block_index);
            flt = (u1 / 50u) & 0xffffu;
          }
          // PrepareSpectrum (:455-473)
          // (flt <= 16384 and the spectrum is int16: the products >> 14 stay within int16, casts not needed)
          re[j] = (re[j] * (int)flt) >> 14;
          im[j] = (im[j] * (int)flt) >> 14;
          if (mine) {
            rb[j].z = norm1 > 5 ? noise[j] << (norm1 - 5) : noise[j] >> (5 - norm1);
            rb[j].w = magn[j];
            RB[k] = rb[j];
            reinterpret_cast<uint16_t*>(&RA[k])[7] = (uint16_t)flt;
            buf[k] = fx_pack(re[j], fx_s16(-im[j]));  // freq_buf[2k], freq_buf[2k+1]
            if (NB > 1 && !nyq && k >= HALF - (HALF >> 2) && k < HALF) {
              hb_psum += nonspeech[j];
              hb_gsum += flt;
            }
          }
        }
      }
      Hw[kX_prevQNoise] = q_noise_new;
      Hw[kX_prevQMagn] = q_magn;
      __syncwarp();

      NSX_PHASE_SYNC();
      // ---- inverse FFT (real_fft.c:74-102): conjugate-symmetric extension,
      // bit reversal, radix-2 stages with data-dependent scaling
      {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int pos = 8 * lane + r;
          int src;
          if (ANA == 256) src = (int)(__brev((unsigned)pos) >> 24);
          else src = (int)(__brev((unsigned)pos) >> 25);
          vr[r] = vi[r] = 0;
          if (act) {
            // (the mirrored half is the conjugate: real_fft.c:86-91)
            const uint32_t m = buf[src <= HALF ? src : ANA - src];
            vr[r] = fx_lo(m);
            vi[r] = src <= HALF ? fx_hi(m) : fx_s16(-fx_hi(m));
          }
        }
      }
      __syncwarp();
      const int out_cifft = fx_warp_cfft<true, ANA>(vr, vi, scr, s_tw, lane);

      // ---- Denormalize (:476), gain (:1462-1496), SynthesisUpdate (:490)
      int y[8];
      int max_o16 = -1;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        y[r] = act ? fx_sat16(fx_shift_w32(vr[r], out_cifft - norm_data)) : 0;
        const int a16 = y[r] > 0 ? y[r] : fx_s16(-y[r]);
        if (act) max_o16 = a16 > max_o16 ? a16 : max_o16;
      }
      int gain = 8192;
      if (gain_map == 1 && block_index > 200 && energy_in > 0) {
        max_o16 = warp_max_i(max_o16);
        int scale_out;
        {
          const int nbits = ANA == 256 ? 9 : 8;
          if (max_o16 == 0) {
            scale_out = 0;
          } else {
            const int t = fx_norm_w32(max_o16 * max_o16);
            scale_out = t > nbits ? 0 : nbits - t;
          }
        }
        int e_out = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) e_out += (y[r] * y[r]) >> scale_out;
        e_out = warp_sum_i(e_out);
        int e_in = energy_in;
        if (scale_out == 0 && !(e_out & 0x7f800000)) {
          e_out = fx_shift_w32(e_out, 8 + scale_out - scale_in);
        } else {
          e_in >>= 8 + scale_out - scale_in;
        }
        int ratio = fx_s16((e_out + e_in / 2) / e_in);
        ratio = ratio > 256 ? 256 : (ratio < 0 ? 0 : ratio);
        const int g1 = T->factor1[ratio];
        const int g2 = T->factor2[mode > 0 ? mode - 1 : 0][ratio];
        gain = fx_s16(fx_s16((fx_s16(16384 - prior_ns) * g1) >> 14) + fx_s16((prior_ns * g2) >> 14));
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int a = fx_mul_rsft_round(s_win[act ? lane + LANES * r : 0], y[r], 14);   // an int16 already (as above)
        const int t = fx_sat16(fx_mul_rsft_round(a, gain, 13));
        const int prev = r < 3 ? syn_h[r] : 0;
        y[r] = fx_sat16(prev + t);
      }
#pragma unroll
      for (int r = 0; r < 5; ++r) outv[r] = y[r];
#pragma unroll
      for (int r = 0; r < 3; ++r) syn_h[r] = y[r + 5];

      // ---- high-band gain (nsx_core.c:2057-2108)
      if (NB > 1) {
        hb_psum = warp_sum_u(hb_psum) & 0xffffu;
        hb_gsum = warp_sum_u(hb_gsum);
        const int avg_prob = fx_s16(4096 - (int)(hb_psum >> (STAGES - 7)));
        const int avg_gain = fx_s16((int)(hb_gsum >> (STAGES - 3)));
        const int gain_mod = avg_prob < 3607 ? avg_prob : 3607;
        int g;
        if (avg_prob < 2048) {
          g = fx_s16((gain_mod << 1) + (avg_gain >> 1));
        } else {
          g = fx_s16((3 * avg_gain) >> 2);
          g = fx_s16(g + gain_mod);
        }
        g = g > 16384 ? 16384 : (g < fx_s16(denoise_bound) ? fx_s16(denoise_bound) : g);
        hb_gain = g;
        hb_apply = true;
      }
    }

    // ---- outputs
    if (act) {
      int16_t* dst = gout + (size_t)f * (size_t)p.out_frame_stride;
#pragma unroll
      for (int r = 0; r < 5; ++r) dst[lane + LANES * r] = (int16_t)outv[r];
    }
    if (NB > 1) {
      // dataBufHBFX delay lines (nsx_core.c:1581-1592, 2046-2054, 2112-2119)
#pragma unroll
      for (int b = 0; b < NB - 1; ++b) {
        int hb[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) hb[r] = r < 3 ? hb_h[b][r] : cur[b + 1 < NB ? b + 1 : 0][r - 3];
#pragma unroll
        for (int r = 0; r < 3; ++r) hb_h[b][r] = hb[r + 5];
        if (act) {
          int16_t* dst = gout + (size_t)f * (size_t)p.out_frame_stride + (size_t)(b + 1) * (size_t)p.out_band_stride;
#pragma unroll
          for (int r = 0; r < 5; ++r)
            dst[lane + LANES * r] = (int16_t)(hb_apply ? fx_s16((hb_gain * hb[r]) >> 14) : hb[r]);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int r = 0; r < 5; ++r) cur[b][r] = nxt[b][r];
    __syncwarp();
    { int* t = Hr; Hr = Hw; Hw = t; }
  }

  // ---- state: shared / registers -> HBM (bulk copies; the histories go back through the scratch)
  if (act) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      reinterpret_cast<int16_t*>(scr)[lane + LANES * r] = (int16_t)ana_h[r];
      reinterpret_cast<int16_t*>(scr + 48)[lane + LANES * r] = (int16_t)syn_h[r];
#pragma unroll
      for (int b = 0; b < NB - 1; ++b)
        reinterpret_cast<int16_t*>(scr + 96 + 48 * b)[lane + LANES * r] = (int16_t)hb_h[b][r];
    }
  }
  bulk_store_fence();
  __syncwarp();
  if (lane == 0) {
    bulk_store(gS, Hr, kNsxHdrWords * 4);
    bulk_store(gS + kNsxOffAna, scr, kHistBytes);
    bulk_store(gS + kNsxOffSyn, scr + 48, kHistBytes);
#pragma unroll
    for (int b = 0; b < NB - 1; ++b) bulk_store(gS + kNsxOffHb + 48 * b, scr + 96 + 48 * b, kHistBytes);
    bulk_store(gS + kNsxOffRecA, RA, kRecBytes);
    bulk_store(gS + kNsxOffRecB, RB, kRecBytes);
    bulk_store_drain();
  }
}

}  // namespace nsb200

#endif  // AUDIOSIGNALPROCESS_B200_NSX_KERNEL_CUH_
