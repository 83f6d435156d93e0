// Float noise suppressor (WebRtcNs_Analyze + WebRtcNs_Process fused) as one
// sm_100a kernel: one warp per stream walks F 10-ms frames; lanes own bins.
//
// What it replaces in the reference (webrtc/modules/audio_processing/ns/):
//   WebRtcNs_AnalyzeCore   ns_core.c:1043-1181
//   WebRtcNs_ProcessCore   ns_core.c:1183-1415
//   and everything they call: UpdateBuffer :855, Windowing :969, Energy :951,
//   FFT/IFFT :886/:923 (Ooura rdft, utility/fft4g.c:324), NoiseEstimation :217,
//   ComputeSnr :566, ComputeSpectralFlatness :523, ComputeSpectralDifference
//   :595, FeatureUpdate :755, FeatureParameterExtraction :293, SpeechNoiseProb
//   :642, UpdateNoiseEstimate :800, ComputeDdBasedWienerFilter :985.
// All three reference call sites feed Analyze and Process the same frame
// (test_ns_module.cpp:97-99, libapm/src/apm_ns.cpp:69-74), so analyzeBuf ==
// dataBuf and magnPrevAnalyze == magnPrevProcess: one forward FFT per frame.
//
// Data layout
//   HBM  per-stream state slab of kNsfStateWords 32-bit words (SoA inside the
//        slab: header scalars | analysis history | synthesis overlap | HB
//        delay lines | initMagnEst | per-bin records of 12 floats), plus a
//        cold slab of 3x1000 histogram counters per stream.
//   SMEM per warp: header (double buffered), per-bin records, FFT scratch, analysis block
//        [history | frame], synthesis overlap. Per CTA: window, twiddles, log(i) table,
//        mbarriers.  State and tables arrive and leave as TMA bulk copies issued by one lane.
//   REGS the frame in flight: spectrum slots, per-bin statistics, prefetched PCM words.
#ifndef AUDIOSIGNALPROCESS_B200_NSF_KERNEL_CUH_
#define AUDIOSIGNALPROCESS_B200_NSF_KERNEL_CUH_

#include "ns_warp.cuh"
#include "nsf_layout.h"

namespace nsb200 {

// 4-warp CTAs, 4 per SM (128 registers, no spills): 16 resident streams per SM -- the best of 14/16/18/20
// warps per SM; 1/2/4/8 warps per CTA run alike (profiles/README.md), and four share one copy of the
// 6.6 KB table block, which pays for the staging arrays of the sequential sums.
#ifndef NSF_WARPS_PER_CTA
#define NSF_WARPS_PER_CTA 4
#endif
#ifndef NSF_FRAME_SYNC
#define NSF_FRAME_SYNC 0
#endif
constexpr int kNsfWarpsPerCta = NSF_WARPS_PER_CTA;
#ifndef NSF_CTAS_PER_SM
#define NSF_CTAS_PER_SM 4
#endif
constexpr int kNsfCtasPerSm = NSF_CTAS_PER_SM;
// Per-CTA shared memory: the table image (nsf_layout.h: win 256 | tw 512 | logi 132 | pad | regrouped
// twiddles 240) followed by the mbarriers of the TMA bulk copies: one for the tables, two per warp
// (header + sample histories | per-bin records).
static_assert(kNsfImgTw12 + 2 * kFftTw12F2 == kNsfImgSplit && kNsfImgOtw + 2 * kOouraTwF2 <= kNsfTableImgWords, "table image layout");
constexpr int kNsfCtaBarWords = (2 * (1 + 2 * kNsfWarpsPerCta) + 3) / 4 * 4;
constexpr int kNsfCtaTableWords = kNsfTableImgWords + kNsfCtaBarWords;
// Per-warp shared memory (32-bit words): header x2 | bin records | FFT scratch | analysis block
// (256) | synthesis overlap (96) | split: aux arrays + Process block (256) | high-band delay blocks.
// The sample histories live in shared memory, not registers: the register budget of the per-bin
// phases decides how far ptxas can interleave their dependent division chains.
constexpr int kNsfStageWords = 4 * 132;   // four staging arrays of the sequential sums (chain_sum4)
constexpr int kNsfWarpWordsBase = 2 * kNsfHdrWords + 129 * kNsfBinRec + 2 * kFftScratchF2 + 256 + 96 + kNsfStageWords;
template <bool SPLIT, int NB>
struct NsfWarpWords {
  static constexpr int value = kNsfWarpWordsBase + (SPLIT ? 4 * kNsfAuxStride + 256 : 0) + (NB - 1) * 256;
};
constexpr int kNsfWarpWordsMax = NsfWarpWords<true, 3>::value;

template <int ANA>
struct NsfGeo {
  static constexpr int kFrame = ANA == 256 ? 160 : 80;
  static constexpr int kBins = ANA / 2 + 1;
  static constexpr int kNC = ANA / 2;           // complex points
  static constexpr int kL = kNC / 4;            // FFT lanes
  static constexpr int kSlots = kNC / 32 + 1;   // bins per lane incl. Nyquist slot
  static constexpr int kPairs = kNC / 64;       // slot pairs (2g, 2g+1) worked on packed; the Nyquist slot stays scalar
  static constexpr int kFP = kFrame / 2;        // frame sample pairs (80 / 40)
  static constexpr int kHP = (ANA - kFrame) / 2;  // history pairs (48 / 24)
  static constexpr int kTP = ANA / 2;           // all pairs
  static constexpr int kStg = ANA == 256 ? 132 : 68;   // words per staging array of chain_sum4: >= kBins, = 4 (mod 32)
};

// ---- the per-bin phases, written once for one bin (T = float: the Nyquist slot) and for two
// (T = float2: slot pairs, packed arithmetic -- ns_warp.cuh)
#define NSF_PAIR(a, g) make_float2((a)[2 * (g)], (a)[2 * (g) + 1])
#define NSF_UNPAIR(a, g, v) ((a)[2 * (g)] = (v).x, (a)[2 * (g) + 1] = (v).y)

// ComputeSnr (ns_core.c:566-588) and the first loop of SpeechNoiseProb (:660-679), plus the terms of
// covMagnPause, varPause, varMagn (:620-626)
template <class T>
NSB_DEV void nsf_snr_lrt(T magn, T noise, T noisePrev, T magnPrevA, T smoothPrev, T mpause, float avgMagn,
                         float avgPause, T& prevEst, T& logLrt, T& dmdp, T& dpdp, T& dmdm) {
  const T eps = vbcast(0.0001f, T()), one = vbcast(1.f, T());
  prevEst = vmul(vfdiv(magnPrevA, vadd(noisePrev, eps)), smoothPrev);
  const T post = vadd(vfdiv(magn, vadd(noise, eps)), vbcast(-1.f, T()));
  const T snrPost = vsel(vgt(magn, noise), post, vbcast(0.f, T()));
  const T snrPrior = vmmadd(vbcast(0.98f, T()), prevEst, vbcast(1.f - 0.98f, T()), snrPost);
  const T sp2 = vadd(snrPrior, snrPrior);   // 2 * snrPrior
  const T t1 = vadd(one, sp2);
  const T t2 = vfdiv(sp2, vadd(t1, eps));
  // besselTmp - log(t1): a product into a sum, so through vmadd; the product by 0.5 after it is exact --
  // contracted into the last sum or not, the same bits
  const T d = vsub(vmadd(vadd(snrPost, one), t2, vneg(vlog_rn(t1))), logLrt);
  logLrt = vadd(logLrt, vmul(vbcast(0.5f, T()), d));
  const T dm = vadd(magn, vbcast(-avgMagn, T())), dp = vadd(mpause, vbcast(-avgPause, T()));
  dmdp = vmul(dm, dp);
  dpdp = vmul(dp, dp);
  dmdm = vmul(dm, dm);
}

// UpdateNoiseEstimate for one bin / two bins (ns_core.c:800-846); prevHigh: the previous bin's speech
// probability exceeds 0.2 (false at bin 0).  When the old and the new smoothing factor agree the second
// estimate is the first one, bit for bit, so the reference's branch is a minimum.
template <class T, class M>
NSB_DEV void nsf_noise_update(T ps, M prevHigh, T magn, T noisePrev, T& mpause, T& noise) {
  const T one = vbcast(1.f, T()), thr = vbcast(0.2f, T());
  const T hi = vbcast(0.99f, T()), lo = vbcast(0.9f, T()), chi = vbcast(1.f - 0.99f, T()), clo = vbcast(1.f - 0.9f, T());
  const T pn = vsub(one, ps);
  const T mix = vmmadd(pn, magn, ps, noisePrev);
  const T nTmp = vmmadd(vsel(prevHigh, hi, lo), noisePrev, vsel(prevHigh, chi, clo), mix);
  const M high = vgt(ps, thr);
  const T nz2 = vmmadd(vsel(high, hi, lo), noisePrev, vsel(high, chi, clo), mix);
  mpause = vsel(vlt(ps, thr), vmadd(vbcast(0.05f, T()), vsub(magn, mpause), mpause), mpause);
  noise = vsel(vlt(nTmp, nz2), nTmp, nz2);
}

// ComputeDdBasedWienerFilter with the flooring of ProcessCore (ns_core.c:985-1007, :1268-1274)
template <class T>
NSB_DEV T nsf_wiener_gain(T magn, T noise, T prevEst, float overdrive, float denoiseBound) {
  const T one = vbcast(1.f, T()), db = vbcast(denoiseBound, T());
  const T post = vadd(vfdiv(magn, vadd(noise, vbcast(0.0001f, T()))), vbcast(-1.f, T()));
  const T cur_est = vsel(vgt(magn, noise), post, vbcast(0.f, T()));
  const T sp = vmmadd(vbcast(0.98f, T()), prevEst, vbcast(1.f - 0.98f, T()), cur_est);
  T g = vfdiv(sp, vadd(vbcast(overdrive, T()), sp));
  g = vsel(vlt(g, db), db, g);
  return vsel(vgt(g, one), one, g);
}

// One bin's (T = float) or two bins' (T = float2, packed arithmetic) three quantile trackers, ns_core.c:236-259.
// The step carries its sign through the division (round-to-nearest is symmetric), so the update is one addition.
// `chain` advances `adv` steps after each tracker: the sequential sums of pass A need none of this.
template <class T, class Chain>
NSB_DEV void nsf_tracker_update(const T lm, T (&lq)[3], T (&dn)[3], const float (&c1)[3], const float (&rc1)[3],
                                const float (&cf)[3], Chain& chain, int adv) {
  const T one = vbcast(1.f, T()), forty = vbcast(40.f, T());
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const T vc1 = vbcast(c1[s], T()), vrc1 = vbcast(rc1[s], T());
    const T delta = vsel(vgt(dn[s], one), vfdiv(forty, dn[s]), forty);
    // one division: QUANTILE*delta/(c+1) upwards, (1-QUANTILE)*delta/(c+1) downwards
    const T w = vsel(vgt(lm, lq[s]), vbcast(0.25f, T()), vbcast(-(1.f - 0.25f), T()));
    lq[s] = vadd(lq[s], vfdiv_r(vmul(w, delta), vc1, vrc1));
    const T upd = vfdiv_r(vmadd(vbcast(cf[s], T()), dn[s], vbcast(1.f / (2.f * 0.01f), T())), vc1, vrc1);
    dn[s] = vsel(vlt(vabs(vsub(lm, lq[s])), vbcast(0.01f, T())), upd, dn[s]);
    chain.advance(adv);
  }
}

// The reference's real-input split seen from one bin (fft4g.c:1234-1256): zk, zm the complex transform at
// k and N/2 - k, w = s_split[k]:  (xd, xs) = (zk.x - zm.x, zk.y + zm.y),
// re = zk.x - (w.x xd - w.y xs), im = zk.y - (w.x xs + w.y xd) -- on (re, im) pairs: five packed operations.
NSB_DEV float2 nsf_real_split(float2 zk, float2 zm, float2 w) {
  const float2 x = vadd(zk, make_float2(-zm.x, zm.y));
  const float2 in = cmul_parts(w.x, x, w.y, x);   // (w.x xd - w.y xs, w.x xs + w.y xd)
  return vadd(zk, make_float2(-in.x, -in.y));
}

NSB_DEV int pad_idx(int k) { return k + ((k >> 6) << 1); }

NSB_DEV float sat_s16f(float v) {  // WEBRTC_SPL_SAT(32767, v, -32768)
  return v > 32767.f ? 32767.f : (v < -32768.f ? -32768.f : v);
}
NSB_DEV int round_s16(float v) {  // FloatS16ToS16, common_audio/include/audio_util.h:41-49
  // v > 0 ? (v >= 32766.5 ? 32767 : (int)(v + .5)) : (v <= -32767.5 ? -32768 : (int)(v - .5));
  // v -+ 0.5 is exact below 2^15, so truncating first and clamping after gives the same integers.
#ifdef __CUDA_ARCH__
  // float -> s16 conversion truncates and clamps in one instruction (checked by the device self-test)
  short r;
  asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(r) : "f"(v + copysignf(0.5f, v)));
  return (int)r;
#else
  const int r = (int)(v + copysignf(0.5f, v));
  return r > 32767 ? 32767 : (r < -32768 ? -32768 : r);
#endif
}

// One 32-bit-lane PCM word as it sits in memory: two int16 samples, or a float pair.
template <bool I16> struct PcmWordT { typedef float2 type; };
template <> struct PcmWordT<true> { typedef uint32_t type; };
template <bool I16>
NSB_DEV void pcm_load(float2& dst, const void* base, size_t off, int w) {
  dst = reinterpret_cast<const float2*>(static_cast<const float*>(base) + off)[w];
}
template <bool I16>
NSB_DEV void pcm_load(uint32_t& dst, const void* base, size_t off, int w) {
  dst = reinterpret_cast<const uint32_t*>(static_cast<const int16_t*>(base) + off)[w];
}
NSB_DEV float2 pcm_unpack(float2 v) { return v; }
NSB_DEV float2 pcm_unpack(uint32_t v) {
  return make_float2((float)(int16_t)(v & 0xffffu), (float)(int16_t)(v >> 16));
}

// ---- threshold re-estimation every 500 frames (ns_core.c:337-517), warp parallel.
// Peak search semantics of the sequential scan (strict '>' both times):
//   peak1 = first occurrence of the maximum;
//   peak2 = first occurrence of the maximum over all other bins, if > 0.
NSB_DEV void nsf_hist_two_peaks(const int* h, int lane, int& w1, int& i1, int& w2, int& i2) {
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
  w1 = bv; i1 = bi;
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
  w2 = cv; i2 = ci;
}

NSB_DEV void nsf_extract_params(float* H, int* hist, int lane) {
  int* hLrt = hist;
  int* hFlat = hist + 1000;
  int* hDiff = hist + 2000;
  // LRT fluctuation (ns_core.c:340-358).  The reference adds the bins' terms in turn in single precision, and
  // `fluct < 0.05` below switches a feature on or off: every lane walks all 1000 bins in that order (once per
  // 500 frames), the counters fetched 32 at a time and handed round by shuffle.
  float avg = 0.f, avgC = 0.f, avgSq = 0.f;
  int num = 0;
  for (int base = 0; base < 1000; base += 32) {
    const int mine = base + lane < 1000 ? __ldcg(hLrt + base + lane) : 0;  // counters are bumped by L2 atomics: bypass L1
    const int nl = 1000 - base < 32 ? 1000 - base : 32;
    for (int l = 0; l < nl; ++l) {
      const int c = __shfl_sync(kFullMask, mine, l);
      const float mid = ((float)(base + l) + 0.5f) * 0.1f;
      if (mid <= 1.f) { avg += c * mid; num += c; }
      avgSq += c * mid * mid;
      avgC += c * mid;
    }
  }
  if (num > 0) avg = avg / (float)num;
  avgC = avgC / 500.f;
  avgSq = avgSq / 500.f;
  const float fluct = avgSq - avg * avgC;
  float thrLrt;
  if (fluct < 0.05f) {
    thrLrt = 1.f;
  } else {
    thrLrt = 1.2f * avg;
    if (thrLrt < 0.2f) thrLrt = 0.2f;
    if (thrLrt > 1.f) thrLrt = 1.f;
  }
  H[kH_priorPars + 0] = thrLrt;

  int w1, i1, w2, i2;
  // spectral flatness: bin 0.05
  nsf_hist_two_peaks(hFlat, lane, w1, i1, w2, i2);
  float p1 = w1 > 0 ? ((float)i1 + 0.5f) * 0.05f : 0.f;
  float p2 = w2 > 0 ? ((float)i2 + 0.5f) * 0.05f : 0.f;
  int useFlat = 1;
  if (fabsf(p2 - p1) < 2 * 0.05f && (float)w2 > 0.5f * (float)w1) {
    w1 += w2;
    p1 = 0.5f * (p1 + p2);
  }
  if (w1 < 150 || p1 < 0.6f) useFlat = 0;
  if (useFlat) {
    float t = 0.9f * p1;
    if (t < 0.1f) t = 0.1f;
    if (t > 0.95f) t = 0.95f;
    H[kH_priorPars + 1] = t;
  }
  // spectral difference: bin 0.1
  nsf_hist_two_peaks(hDiff, lane, w1, i1, w2, i2);
  p1 = w1 > 0 ? ((float)i1 + 0.5f) * 0.1f : 0.f;
  p2 = w2 > 0 ? ((float)i2 + 0.5f) * 0.1f : 0.f;
  int useDiff = 1;
  if (fabsf(p2 - p1) < 2 * 0.1f && (float)w2 > 0.5f * (float)w1) {
    w1 += w2;
    p1 = 0.5f * (p1 + p2);
  }
  float td = 1.2f * p1;
  if (w1 < 150) useDiff = 0;
  if (td < 0.16f) td = 0.16f;
  if (td > 1.f) td = 1.f;
  H[kH_priorPars + 3] = td;
  if (fluct < 0.05f) useDiff = 0;
  const float fsum = (float)(1 + useFlat + useDiff);
  H[kH_priorPars + 4] = 1.f / fsum;
  H[kH_priorPars + 5] = (float)useFlat / fsum;
  H[kH_priorPars + 6] = (float)useDiff / fsum;
  __syncwarp();
  for (int i = lane; i < 3000; i += 32) hist[i] = 0;
}

// ---- the kernel -------------------------------------------------------------
// ANA: 256 (16/32/48 kHz band 0) or 128 (8 kHz). NB: number of bands (1..3).
// I16: PCM is int16 (rounded on output like IFChannelBuffer::RefreshI,
// channel_buffer.cc:55-60) or float in int16 scale (the WebRtcNs_Process ABI).
// SPLIT: Analyze and Process are fed different band-0 signals (p.ana_in / p.in), as when an
// echo canceller sits between them (audio_processing_impl.cc:625-631).  The analysis keeps its
// own history, the process side gets its own forward FFT, and what one frame implies in the
// fused case -- noise == noisePrev, magnPrevAnalyze == magnPrevProcess, speechProb and
// parametricNoise living in registers -- becomes four extra per-bin arrays (nsf_layout.h).
template <int ANA, int NB, bool I16, bool SPLIT>
__global__ void __launch_bounds__(kNsfWarpsPerCta * 32, kNsfCtasPerSm)
nsf_process_kernel(const NsfLaunch p) {
  typedef NsfGeo<ANA> G;
  extern __shared__ float4 nsf_smem4[];
  float* smem = reinterpret_cast<float*>(nsf_smem4);
  float* s_win = smem + kNsfImgWin;
  float2* s_tw = reinterpret_cast<float2*>(smem + kNsfImgTw);
  float* s_logi = smem + kNsfImgLogi;
  float2* s_tw12 = reinterpret_cast<float2*>(smem + kNsfImgTw12);
  const float2* s_split = reinterpret_cast<const float2*>(smem + kNsfImgSplit);   // real-input split of the forward FFT
  const float2* s_otw = reinterpret_cast<const float2*>(smem + kNsfImgOtw);       // its twiddles per (pass, lane)
  mbar_t* bars = reinterpret_cast<mbar_t*>(smem + kNsfTableImgWords);

  const int lane = lane_id();
  // (through a shuffle: the compiler then knows the warp index -- and every shared-memory address and
  // slab pointer derived from it -- to be warp-uniform, keeps them in uniform registers and feeds
  // the bulk-copy instructions from there instead of electing a lane and looping per operand:
  // ~14 -> ~4 instructions per copy, a dozen copies per stream and launch)
  const int warp = __shfl_sync(kFullMask, (int)(threadIdx.x >> 5), 0);
  const NsfTables* T = p.tables;
  const int sidx = (int)blockIdx.x * kNsfWarpsPerCta + warp;
  const bool live = sidx < p.n_streams;

  float* W = smem + kNsfCtaTableWords + warp * NsfWarpWords<SPLIT, NB>::value;
  // Header scalars are double buffered: within a frame every lane reads the
  // frame-start copy (Hr) and writes the next copy (Hw) with warp-uniform
  // values, so there is never a read-modify-write race between lanes.
  float* Hr = W;
  float* Hw = W + kNsfHdrWords;
  float* B = W + 2 * kNsfHdrWords;            // per-bin records
  float2* scr = reinterpret_cast<float2*>(B + 129 * kNsfBinRec);  // FFT scratch (8-byte aligned)
  float2* blkA = scr + kFftScratchF2;          // analysis block [history | frame], 128 pairs (analyzeBuf)
  float2* ovl = blkA + 128;                    // synthesis overlap, 48 pairs (head of syntBuf)
  float* stg = reinterpret_cast<float*>(ovl + 48);                  // staging arrays of the sequential sums (16-byte aligned)
  float* X_noise = stg + kNsfStageWords;                            // split mode: self->noise
  float* X_prob = X_noise + kNsfAuxStride;                          //   self->speechProb
  float* X_param = X_prob + kNsfAuxStride;                          //   self->parametricNoise
  float* X_magnP = X_param + kNsfAuxStride;                         //   self->magnPrevProcess
  float2* blkP = reinterpret_cast<float2*>(X_magnP + kNsfAuxStride);  // split mode: Process block (dataBuf)
  float2* blkH = reinterpret_cast<float2*>(stg + kNsfStageWords + (SPLIT ? 4 * kNsfAuxStride + 256 : 0));  // high-band delay blocks (dataBufHB)

  // ---- tables and state: HBM -> shared by TMA bulk copies.  Everything a stream needs is in
  // flight after a handful of instructions of one lane and one dependent load (the slot), where
  // the table loops, the header, the histories and the records used to be a chain of five
  // dependent global-memory latencies per warp -- most of a launch that walks a single frame.
  mbar_t* barT = bars;                 // tables (per CTA)
  mbar_t* barH = bars + 1 + 2 * warp;  // header + sample histories: needed by the first window / FFT
  mbar_t* barB = barH + 1;             // per-bin records (+ split arrays): awaited after the first FFT
  if (threadIdx.x == 0) mbar_init(barT, 1);
  if (lane == 0) {
    mbar_init(barH, 1);
    mbar_init(barB, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect(barT, sizeof(float) * kNsfTableImgWords);
    bulk_load(smem, T->img[ANA == 256 ? 0 : 1], sizeof(float) * kNsfTableImgWords, barT);
  }
  const int slot = __shfl_sync(kFullMask, live ? (p.slots ? p.slots[sidx] : p.slot_base + sidx) : 0, 0);
  float* gS = p.state + (size_t)slot * kNsfStateWords;
  int* gHist = p.hist + (size_t)slot * kNsfHistWords;
  float* gInitMagn = gS + kNsfOffInitMagn;
  constexpr unsigned kHistBytes = G::kHP * sizeof(float2);             // 384 / 192
  constexpr unsigned kBinBytes = G::kBins * kNsfBinRec * sizeof(float);  // 6192 / 3120
  if (live && lane == 0) {
    // sample histories: the tail of each block is what the previous frame left behind
    mbar_arrive_expect(barH, kNsfHdrWords * 4 + (2 + (SPLIT ? 1 : 0) + (NB - 1)) * kHistBytes);
    bulk_load(Hr, gS, kNsfHdrWords * 4, barH);
    bulk_load(blkA + G::kFP, gS + kNsfOffXHist, kHistBytes, barH);
    bulk_load(ovl, gS + kNsfOffSynt, kHistBytes, barH);
    if (SPLIT) bulk_load(blkP + G::kFP, gS + kNsfOffPHist, kHistBytes, barH);
#pragma unroll
    for (int b = 0; b < NB - 1; ++b) bulk_load(blkH + 128 * b + G::kFP, gS + kNsfOffHb + 96 * b, kHistBytes, barH);
    mbar_arrive_expect(barB, kBinBytes + (SPLIT ? 4 * kNsfAuxStride * 4 : 0));
    bulk_load(B, gS + kNsfOffBins, kBinBytes, barB);
    if (SPLIT) bulk_load(X_noise, gS + kNsfOffAux, 4 * kNsfAuxStride * 4, barB);
  }
  bool state_ready = false;
  // the staging arrays' tails beyond the last bin stay +0.f for the whole launch (chain_sum4 adds them)
  if (lane < 4 * (G::kStg - G::kBins)) stg[(lane / (G::kStg - G::kBins)) * G::kStg + G::kBins + lane % (G::kStg - G::kBins)] = 0.f;
  // Prefetch across CTAs: the stream that the block scheduler will start in this warp's place about
  // one wave from now (batch entry sidx + resident warps) gets its state and first frame pulled
  // into L2, so that its own bulk copies find them there instead of in DRAM.  One wave of state is
  // 17 MB of the 126 MB L2.
  // (prefetch_ahead < 0: a one-frame tick -- the entries past the end of the batch wrap around to its
  // head, whose state this launch has already written back: the next tick's first wave starts warm)
  const int ahead = p.prefetch_ahead < 0 ? -p.prefetch_ahead : p.prefetch_ahead;
  const bool wrap = p.prefetch_ahead < 0 && sidx + ahead >= p.n_streams && sidx + ahead - p.n_streams < sidx;
  if (live && ahead > 0 && (sidx + ahead < p.n_streams || wrap)) {
    const int an = wrap ? sidx + ahead - p.n_streams : sidx + ahead;
    const float* aS = p.state + (size_t)(p.slots ? p.slots[an] : p.slot_base + an) * kNsfStateWords;
    if (lane == 0) {
      bulk_prefetch_l2(aS, (kNsfOffHb + 96 * (NB - 1)) * 4);   // header | histories | overlap | high-band delay
      bulk_prefetch_l2(aS + kNsfOffBins, kBinBytes);
      if (SPLIT) bulk_prefetch_l2(aS + kNsfOffSplit, (96 + 4 * kNsfAuxStride) * 4);
    }
    if (p.frames > 0 && !wrap) {   // (the next tick's PCM is not this launch's to know)
      const char* a = static_cast<const char*>(p.in) + (size_t)an * (size_t)p.in_stream_stride * (I16 ? 2 : 4);
      const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(a) & 127u);
      if (lane * 128u < mis + G::kFrame * (I16 ? 2u : 4u)) line_prefetch_l2(a - mis + 128 * lane);
    }
  }

  // PCM addressing: frame pair w of band b.
  const size_t in_base = (size_t)sidx * (size_t)p.in_stream_stride;
  const size_t out_base = (size_t)sidx * (size_t)p.out_stream_stride;
  constexpr int kU = (G::kFP + 31) / 32;  // words per lane per band-frame (3 / 2)

  // Raw frame words (two int16 samples, or a float pair), prefetched one frame ahead and
  // converted where they are used: converting at the load would wait for it on the spot.
  typedef typename PcmWordT<I16>::type PcmWord;
  PcmWord cur[NB][kU];
  auto load_frame = [&](int f, PcmWord (&dst)[NB][kU]) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const size_t off = in_base + (size_t)f * (size_t)p.in_frame_stride + (size_t)b * (size_t)p.in_band_stride;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int w = lane + 32 * u;
        dst[b][u] = PcmWord();
        if (w < G::kFP) pcm_load<I16>(dst[b][u], p.in, off, w);
      }
    }
  };
  auto store_pair = [&](int f, int b, int w, float2 v) {
    const size_t off = out_base + (size_t)f * (size_t)p.out_frame_stride + (size_t)b * (size_t)p.out_band_stride;
    if (I16) {
      const uint32_t lo = (uint32_t)round_s16(v.x) & 0xffffu, hi = (uint32_t)round_s16(v.y) & 0xffffu;
      reinterpret_cast<uint32_t*>(static_cast<int16_t*>(p.out) + off)[w] = lo | (hi << 16);
    } else {
      reinterpret_cast<float2*>(static_cast<float*>(p.out) + off)[w] = v;
    }
  };

  // split mode: band-0 frame of the Analyze signal
  PcmWord curA[kU];
  auto load_ana = [&](int f, PcmWord (&dst)[kU]) {
    const size_t off = (size_t)sidx * (size_t)p.ana_stream_stride + (size_t)f * (size_t)p.ana_frame_stride;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int w = lane + 32 * u;
      dst[u] = PcmWord();
      if (w < G::kFP) pcm_load<I16>(dst[u], p.ana_in, off, w);
    }
  };
  // the first frame's PCM goes out with the state copies, before anything is waited for
  // p.phase (split kernels only): 1 = the Analyze half of every frame alone, 2 = the Process half alone -- the
  // single-stream WebRtcNs_Analyze / WebRtcNs_Process pair, whose statistics must be current in between
  // (WebRtcNs_prior_speech_probability, noise_suppression.c:57-66); 0 = both.
  const bool doA = !SPLIT || p.phase != 2, doP = !SPLIT || p.phase != 1;
  if (doP && live && p.frames > 0) load_frame(0, cur);
  if (SPLIT && doA && live && p.frames > 0) load_ana(0, curA);

  mbar_wait_cta(barT, 0);
  if (!live) return;  // whole warp leaves; no block barriers below
  mbar_wait_warp(barH, 0);
  if (SPLIT) {
    mbar_wait_warp(barB, 0);
    state_ready = true;
    if (reinterpret_cast<const int*>(Hr)[kH_splitValid] == 0) {
      // the stream was driven by the fused kernel so far: one frame for both sides means
      // noise == noisePrev, magnPrevProcess == magnPrevAnalyze, dataBuf == analyzeBuf
      for (int k = lane; k < G::kBins; k += 32) {
        X_noise[k] = B[k * kNsfBinRec + kB_noisePrev];
        X_magnP[k] = B[k * kNsfBinRec + kB_magnPrev];
        X_prob[k] = 0.f;
        X_param[k] = 0.f;
      }
      for (int pr = lane; pr < G::kHP; pr += 32) blkP[G::kFP + pr] = blkA[G::kFP + pr];
      __syncwarp();
    }
  }

  const float overdrive = Hr[kH_overdrive];
  const float denoiseBound = Hr[kH_denoiseBound];
  const int gainmap = reinterpret_cast<const int*>(Hr)[kH_gainmap];
  constexpr float kMagnLenF = (float)G::kBins;
  const float magnLenF = kMagnLenF;

  // optional lock step of the CTA's warps (full CTAs only): warps in the same phase share
  // instruction-cache lines
  const bool cta_sync = NSF_FRAME_SYNC && ((int)blockIdx.x + 1) * kNsfWarpsPerCta <= p.n_streams;
  for (int f = 0; f < p.frames; ++f) {
    if (cta_sync) __syncthreads();
    PcmWord nxt[NB][kU] = {};
    PcmWord nxtA[kU] = {};

    const int* HIr = reinterpret_cast<const int*>(Hr);
    int* HIw = reinterpret_cast<int*>(Hw);
    Hw[lane] = Hr[lane];  // ordered before any Hw update by the __syncwarp below

    // ---- (a) UpdateBuffer: history | new frame -> scratch as sample pairs
    // (all input was read into registers above, so out may alias in: ns_core.c:1225,1357)
    float2 v[4];
    {
      float2 t[2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (lane + 32 * u < G::kHP) t[u] = blkA[G::kFP + lane + 32 * u];
      __syncwarp();
      if (doA) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (lane + 32 * u < G::kHP) blkA[lane + 32 * u] = t[u];
#pragma unroll
        for (int u = 0; u < kU; ++u)
          if (lane + 32 * u < G::kFP) blkA[G::kHP + lane + 32 * u] = pcm_unpack(SPLIT ? curA[u] : cur[0][u]);
      }
      __syncwarp();
      // The next frame's loads are issued only now, after this frame's words were consumed: the
      // consumer waits on a scoreboard shared by every load this static instruction has in
      // flight, so issuing the prefetch first made it wait for the prefetch itself (10 % of the
      // kernel's stall samples sat on the int16 unpack above).
      if (f + 1 < p.frames) {
        if (doP) load_frame(f + 1, nxt);
        if (SPLIT && doA) load_ana(f + 1, nxtA);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = lane < G::kL ? blkA[lane + G::kL * j] : make_float2(0.f, 0.f);
    }

    // ---- (b) Windowing + Energy (ns_core.c:1070-1071 / 1237-1238)
    float energy1 = 0.f;
    if (lane < G::kL) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 w = reinterpret_cast<const float2*>(s_win)[lane + G::kL * j];
        v[j] = vmul(v[j], w);
        const float2 sq = vmul(v[j], v[j]);
        energy1 += sq.x + sq.y;
      }
    }
    // Energy (ns_core.c:951) is a sum of squares: zero iff every term is, so the zero-input test is
    // one vote, and the five dependent shuffle + add steps of the sum itself (needed only by the
    // gain compensation much later) are issued next to the FFT, where there is independent work to
    // hide them behind.  The reductions are a tenth of this kernel's stall samples.
    const bool nzA = doA && __any_sync(kFullMask, energy1 != 0.f);
    bool nzP = nzA;   // the same frame on the Process side unless SPLIT

    float2 o0[kU];  // band-0 output pairs
    float hbGain = 1.f;
    bool hbApplyGain = false;

    // values handed from the analysis to the process part (fused: in registers; split: the
    // process part reloads them from the state it shares with Analyze)
    int blockInd = HIr[kH_blockInd];
    float magn[G::kSlots];   // the spectrum itself waits in scratch (scr[k]) for the Wiener gain
    float noise[G::kSlots], prevEst[G::kSlots], parametric[G::kSlots], prob[G::kSlots];
    float noisePrev[G::kSlots], logLrt[G::kSlots], mpause[G::kSlots];
    float prior = 0.f;

    if (nzA) {   // energy of the analysed frame != 0 (ns_core.c:1071)
      // ======== WebRtcNs_AnalyzeCore (zero input: no statistics update, ns_core.c:1072-1082)
      blockInd = HIr[kH_blockInd] + 1;
      HIw[kH_blockInd] = blockInd;
      const int updateParsFlag = HIr[kH_modelUpd0];

      // ---- (c) forward FFT (ns_core.c:886-911)
      if (!SPLIT) energy1 = warp_sum(energy1);
      ooura_fwd<G::kNC>(v, scr, s_otw, lane);
      if (lane < G::kL) {
#pragma unroll
        for (int q = 0; q < 4; ++q) scr[pad_idx(ooura_out_index<G::kNC>(lane, q))] = v[q];
      }
      __syncwarp();

      if (!state_ready) {   // the per-bin records: first frame of a launch only
        mbar_wait_warp(barB, 0);
        state_ready = true;
      }
      float lmagn[G::kSlots];
      float smoothPrev[G::kSlots], magnPrevA[G::kSlots];
      {
        float re[G::kSlots], im[G::kSlots];
#pragma unroll
        for (int j = 0; j < G::kSlots; ++j) {
          const bool nyq = (j == G::kSlots - 1);
          const int k = nyq ? G::kNC : lane + 32 * j;
          const float2 zk = scr[pad_idx(k & (G::kNC - 1))];
          const float2 zm = scr[pad_idx((G::kNC - k) & (G::kNC - 1))];
          // the reference's real-input split (fft4g.c:1234-1256 and :352-354) seen from bin k: with
          // (wr, wi) = s_split[k] -- wi negated on the mirror side, (0, +-1/2) at bins 0 and N/2 so that
          // the same expression yields a[0] +- a[1] -- the same additions and products, rounding for rounding
          const float2 w = s_split[k];
          const float2 X = nsf_real_split(zk, zm, w);
          re[j] = X.x;
          im[j] = (nyq || k == 0) ? 0.f : X.y;
          const float4 r2 = *reinterpret_cast<const float4*>(B + k * kNsfBinRec + 8);  // noisePrev magnPrev logLrt pause
          noisePrev[j] = r2.x;
          magnPrevA[j] = r2.y;
          logLrt[j] = r2.z;
          mpause[j] = r2.w;
          if (!nyq || lane == 0) stg[3 * G::kStg + k] = r2.w;   // avgPause       ns_core.c:609
        }
        // magnitudes, their logarithms (float)log((double)magn) (ns_core.c:228) and the terms of the other
        // three sums of pass A (below), bin by bin
#pragma unroll
        for (int g = 0; g < G::kPairs; ++g) {
          const float2 re2 = NSF_PAIR(re, g), im2 = NSF_PAIR(im, g);
          const float2 e = vmmadd(re2, re2, im2, im2);
          const float2 m = vsqrt_p1(e);
          const float2 lm = vlog_rn(m);
          NSF_UNPAIR(magn, g, m);
          NSF_UNPAIR(lmagn, g, lm);
          const int k = lane + 64 * g;
          stg[k] = e.x;                                   // signalEnergy   :1090
          stg[k + 32] = e.y;
          stg[G::kStg + k] = m.x;                         // sumMagn        :1091
          stg[G::kStg + k + 32] = m.y;
          stg[2 * G::kStg + k] = k == 0 ? 0.f : lm.x;     // flatness numerator, bins >= 1   :540-542
          stg[2 * G::kStg + k + 32] = lm.y;
        }
        {
          constexpr int j = G::kSlots - 1;
          const float e = re[j] * re[j] + im[j] * im[j];
          magn[j] = nsb_sqrtf_p1(e);
          lmagn[j] = nsb_log_rn(magn[j]);
          if (lane == 0) {
            stg[G::kNC] = e;
            stg[G::kStg + G::kNC] = magn[j];
            stg[2 * G::kStg + G::kNC] = lmagn[j];
          }
        }
        __syncwarp();  // every lane has read its mirrored points: scratch is free again
        if (!SPLIT) {
#pragma unroll
          for (int j = 0; j < G::kSlots; ++j) {
            const bool nyq = (j == G::kSlots - 1);
            if (!nyq || lane == 0) scr[nyq ? G::kNC : lane + 32 * j] = make_float2(re[j], im[j]);
          }
        }
      }
      // ---- pass A: signalEnergy, sumMagn, sum of lmagn, sum of magnAvgPause, each summed bin after bin as the
      // reference sums them (ChainSum4: four lanes, one chain each).  A 129-step dependent chain: it advances
      // between the tracker updates of (d), which need none of its results.
      ChainSum4<G::kStg, G::kStg> chainA(stg, lane);

      // ---- (d) NoiseEstimation (ns_core.c:217-285)
      int updates = HIr[kH_updates];
      if (updates < 200) updates++;
      HIw[kH_updates] = updates;
      int cnt[3];
      float c1[3], cf[3], rc1[3];   // rc1: reciprocal of counter+1, shared by every bin's two divisions
      // which tracker (if any) is latched into `quantile` this frame: the last one whose
      // counter expired once updates >= 200, else tracker 2 during start-up (:262-280)
      int sel = updates < 200 ? 2 : -1;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        cnt[s] = HIr[kH_counter + s];
        const bool latch = cnt[s] >= 200;
        if (latch && updates >= 200) sel = s;
        HIw[kH_counter + s] = (latch ? 0 : cnt[s]) + 1;
        c1[s] = (float)(cnt[s] + 1);
        rc1[s] = frcp_nr(c1[s]);
        cf[s] = (float)cnt[s];
      }
      // slot pairs (bins lane + 64 g and lane + 64 g + 32) run packed, the Nyquist slot scalar
#pragma unroll
      for (int g = 0; g < G::kPairs; ++g) {
        float* Ra = B + (lane + 64 * g) * kNsfBinRec;
        float* Rb = Ra + 32 * kNsfBinRec;
        const float4 a0 = *reinterpret_cast<float4*>(Ra), a1 = *reinterpret_cast<float4*>(Ra + 4);
        const float4 b0 = *reinterpret_cast<float4*>(Rb), b1 = *reinterpret_cast<float4*>(Rb + 4);
        float2 lq[3] = {make_float2(a0.x, b0.x), make_float2(a0.y, b0.y), make_float2(a0.z, b0.z)};
        float2 dn[3] = {make_float2(a0.w, b0.w), make_float2(a1.x, b1.x), make_float2(a1.y, b1.y)};
        smoothPrev[2 * g] = a1.w;
        smoothPrev[2 * g + 1] = b1.w;
        nsf_tracker_update(make_float2(lmagn[2 * g], lmagn[2 * g + 1]), lq, dn, c1, rc1, cf, chainA, 4);
        noise[2 * g] = a1.z;
        noise[2 * g + 1] = b1.z;
        *reinterpret_cast<float4*>(Ra) = make_float4(lq[0].x, lq[1].x, lq[2].x, dn[0].x);
        // whole 16-byte groups only: scalar accesses at a 12-word lane stride are 4-way bank conflicts
        *reinterpret_cast<float4*>(Ra + 4) = make_float4(dn[1].x, dn[2].x, a1.z, a1.w);
        *reinterpret_cast<float4*>(Rb) = make_float4(lq[0].y, lq[1].y, lq[2].y, dn[0].y);
        *reinterpret_cast<float4*>(Rb + 4) = make_float4(dn[1].y, dn[2].y, b1.z, b1.w);
      }
      {
        float* R = B + G::kNC * kNsfBinRec;
        const float4 r0 = *reinterpret_cast<float4*>(R);      // lq0 lq1 lq2 dens0
        const float4 r1 = *reinterpret_cast<float4*>(R + 4);  // dens1 dens2 quantile smooth
        float lq[3] = {r0.x, r0.y, r0.z};
        float dn[3] = {r0.w, r1.x, r1.y};
        smoothPrev[G::kSlots - 1] = r1.w;
        nsf_tracker_update(lmagn[G::kSlots - 1], lq, dn, c1, rc1, cf, chainA, 4);
        noise[G::kSlots - 1] = r1.z;
        if (lane == 0) {
          *reinterpret_cast<float4*>(R) = make_float4(lq[0], lq[1], lq[2], dn[0]);
          *reinterpret_cast<float4*>(R + 4) = make_float4(dn[1], dn[2], r1.z, r1.w);
        }
      }
      float sigE, sumMagn, sumLog, sumPause;
      {
        const float c = chainA.finish();
        sigE = __shfl_sync(kFullMask, c, 0);
        sumMagn = __shfl_sync(kFullMask, c, 1);
        sumLog = __shfl_sync(kFullMask, c, 2);
        sumPause = __shfl_sync(kFullMask, c, 3);
      }
      const float signalEnergy = NSB_FDIV_C(sigE, kMagnLenF);
      // latch (every frame during start-up, then once per tracker period) outside the slot loop,
      // so that the 15 tracker updates above form one block the scheduler can interleave
      if (sel >= 0) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < G::kSlots; ++j) {
          const bool nyq = (j == G::kSlots - 1);
          float* R = B + (nyq ? G::kNC : lane + 32 * j) * kNsfBinRec;
          const float q = nsb_exp_rn(R[sel]);   // (float)exp(lquantile): ns_core.c:266,277
          noise[j] = q;
          if (!nyq || lane == 0) R[kB_quantile] = q;
        }
      }

      // ---- (e) start-up: white / pink parametric noise (ns_core.c:1088-1162)
      if (blockInd < 50) {
        // The least-squares fit below subtracts nearly equal products of these sums: summed in any other
        // order than the reference's (bins 5, 6, ... in turn, ns_core.c:1088-1100) its result moves by 1e-5
        // relative -- by far the largest non-branching deviation the kernel could have (oracle experiment,
        // profiles/r2_float_parity.md).  Fifty frames per stream: every lane runs the sequential sums.
        float slm = 0.f, slilm = 0.f;
#pragma unroll
        for (int j = 0; j < G::kSlots - 1; ++j) {
#pragma unroll 1
          for (int l = (j == 0 ? 5 : 0); l < 32; ++l) {
            const float lm = __shfl_sync(kFullMask, lmagn[j], l);
            slm += lm;
            slilm += s_logi[l + 32 * j] * lm;
          }
        }
        slm += lmagn[G::kSlots - 1];   // bin N/2 sits in every lane
        slilm += s_logi[G::kNC] * lmagn[G::kSlots - 1];
        const float sli = T->sum_log_i[ANA == 256 ? 0 : 1];
        const float slisq = T->sum_log_i_sq[ANA == 256 ? 0 : 1];
        float white = Hr[kH_white] + sumMagn / magnLenF * overdrive;
        Hw[kH_white] = white;
        float t1 = slisq * (float)(G::kBins - 5);
        t1 -= sli * sli;
        float t2 = slisq * slm - sli * slilm;
        float t3 = t2 / t1;
        if (t3 < 0.f) t3 = 0.f;
        const float pinkNum = Hr[kH_pinkNum] + t3;
        Hw[kH_pinkNum] = pinkNum;
        t2 = sli * slm;
        t2 -= (float)(G::kBins - 5) * slilm;
        t3 = t2 / t1;
        if (t3 < 0.f) t3 = 0.f;
        if (t3 > 1.f) t3 = 1.f;
        const float pinkExp = Hr[kH_pinkExp] + t3;
        Hw[kH_pinkExp] = pinkExp;
        float pnum = 0.f, pexp = 0.f;
        if (pinkExp > 0.f) {
          pnum = nsb_exp_rn(pinkNum / (float)(blockInd + 1));
          pnum *= (float)(blockInd + 1);
          pexp = pinkExp / (float)(blockInd + 1);
        }
#pragma unroll
        for (int j = 0; j < G::kSlots; ++j) {
          const bool nyq = (j == G::kSlots - 1);
          const int k = nyq ? G::kNC : lane + 32 * j;
          if (pinkExp == 0.f) {
            parametric[j] = white;
          } else {
            // opaque to the optimiser: the bin index is loop-invariant, and with it powf's
            // log2(ub) was hoisted out of the frame loop into the kernel prologue -- 200
            // instructions per launch for a branch taken during a stream's first half second
            // (8 % of a launch that walks one frame)
            int kub = k < 5 ? 5 : k;
            asm volatile("" : "+r"(kub));
            const float ub = (float)kub;
            parametric[j] = nsb_div_pow_rn(pnum, ub, T->logk_d[kub], pexp);   // (float)(parametric_num / pow(use_band, parametric_exp))
          }
          noise[j] *= (float)blockInd;
          const float t = parametric[j] * (float)(50 - blockInd);
          noise[j] += t / (float)(blockInd + 1);
          noise[j] /= 50.f;
        }
      }
      // ---- (f) running energy normaliser (ns_core.c:1165-1169)
      float feat5 = Hr[kH_feat + 5];
      if (blockInd < 200) {
        feat5 *= (float)blockInd;
        feat5 += signalEnergy;
        feat5 /= (float)(blockInd + 1);
        Hw[kH_feat + 5] = feat5;
      }

      // ---- (g) ComputeSnr (ns_core.c:566-588) and the first loop of SpeechNoiseProb (:660-679): neither needs
      // anything from FeatureUpdate, and the sum over the updated logLrtTimeAvg rides in pass B
      const float avgPause = NSB_FDIV_C(sumPause, kMagnLenF);
      const float avgMagn = NSB_FDIV_C(sumMagn, kMagnLenF);
#pragma unroll
      for (int g = 0; g < G::kPairs; ++g) {
        float2 pe, ll = NSF_PAIR(logLrt, g), dmdp, dpdp, dmdm;
        nsf_snr_lrt(NSF_PAIR(magn, g), NSF_PAIR(noise, g), NSF_PAIR(noisePrev, g), NSF_PAIR(magnPrevA, g),
                    NSF_PAIR(smoothPrev, g), NSF_PAIR(mpause, g), avgMagn, avgPause, pe, ll, dmdp, dpdp, dmdm);
        NSF_UNPAIR(prevEst, g, pe);
        NSF_UNPAIR(logLrt, g, ll);
        // terms of the four sums of pass B (ComputeSpectralDifference :620-626, SpeechNoiseProb :678)
        const int k = lane + 64 * g;
        stg[k] = dmdp.x;
        stg[k + 32] = dmdp.y;
        stg[G::kStg + k] = dpdp.x;
        stg[G::kStg + k + 32] = dpdp.y;
        stg[2 * G::kStg + k] = dmdm.x;
        stg[2 * G::kStg + k + 32] = dmdm.y;
        stg[3 * G::kStg + k] = ll.x;
        stg[3 * G::kStg + k + 32] = ll.y;
      }
      {
        constexpr int j = G::kSlots - 1;
        float dmdp, dpdp, dmdm;
        nsf_snr_lrt(magn[j], noise[j], noisePrev[j], magnPrevA[j], smoothPrev[j], mpause[j], avgMagn, avgPause,
                    prevEst[j], logLrt[j], dmdp, dpdp, dmdm);
        if (lane == 0) {
          stg[G::kNC] = dmdp;
          stg[G::kStg + G::kNC] = dpdp;
          stg[2 * G::kStg + G::kNC] = dmdm;
          stg[3 * G::kStg + G::kNC] = logLrt[j];
        }
      }
      __syncwarp();
      // ---- pass B: covMagnPause, varPause, varMagn, sum of logLrtTimeAvg, bin after bin -- advancing between the
      // exponentials of SpeechNoiseProb's last loop (ns_core.c:743-747), which need no feature: (float)exp(-logLrt)
      float cov, varP, varM, lsum;
      float invLrt[G::kSlots];
      {
        ChainSum4<G::kStg, G::kStg> chainB(stg, lane);
#pragma unroll
        for (int j = 0; j < G::kSlots; ++j) {
          invLrt[j] = nsb_exp_rn(-logLrt[j]);
          chainB.advance(7);
        }
        const float c = chainB.finish();
        cov = __shfl_sync(kFullMask, c, 0);
        varP = __shfl_sync(kFullMask, c, 1);
        varM = __shfl_sync(kFullMask, c, 2);
        lsum = __shfl_sync(kFullMask, c, 3);
      }

      // ---- (h) FeatureUpdate (ns_core.c:755-791)
      float feat0, feat4;
      {
        // spectral flatness (:523-556); magn >= 1 so the log(0) exit is dead
        float den = sumMagn - __shfl_sync(kFullMask, magn[0], 0);
        den = NSB_FDIV_C(den, kMagnLenF);
        const float num = NSB_FDIV_C(sumLog, kMagnLenF);
        const float sf = fdiv(nsb_exp_rn(num), den);
        feat0 = Hr[kH_feat + 0];
        feat0 += 0.3f * (sf - feat0);
        Hw[kH_feat + 0] = feat0;
        // spectral difference (:595-634)
        cov = NSB_FDIV_C(cov, kMagnLenF);
        varP = NSB_FDIV_C(varP, kMagnLenF);
        varM = NSB_FDIV_C(varM, kMagnLenF);
        float feat6 = Hr[kH_feat + 6] + signalEnergy;
        float ad = varM - fdiv(cov * cov, varP + 0.0001f);
        ad = fdiv(ad, feat5 + 0.0001f);
        feat4 = Hr[kH_feat + 4];
        feat4 += 0.3f * (ad - feat4);
        Hw[kH_feat + 4] = feat4;
        if (updateParsFlag >= 1) {
          int c3 = HIr[kH_modelUpd3] - 1;
          if (c3 > 0) {
            // histogram update (:309-334) -- uses the LRT mean of the previous frame
            if (lane == 0) {
              const float f3 = Hr[kH_feat + 3];
              if (f3 < 1000 * 0.1f && f3 >= 0.f) atomicAdd(gHist + (int)NSB_FDIV_C(f3, 0.1f), 1);
              if (feat0 < 1000 * 0.05f && feat0 >= 0.f) atomicAdd(gHist + 1000 + (int)NSB_FDIV_C(feat0, 0.05f), 1);
              if (feat4 < 1000 * 0.1f && feat4 >= 0.f) atomicAdd(gHist + 2000 + (int)NSB_FDIV_C(feat4, 0.1f), 1);
            }
          }
          if (c3 == 0) {
            __threadfence_block();
            __syncwarp();
            nsf_extract_params(Hw, gHist, lane);
            c3 = 500;
            if (updateParsFlag == 1) {
              HIw[kH_modelUpd0] = 0;
            } else {
              feat6 = feat6 / 500.f;
              feat5 = 0.5f * (feat6 + feat5);
              Hw[kH_feat + 5] = feat5;
              feat6 = 0.f;
            }
          }
          HIw[kH_modelUpd3] = c3;
        }
        Hw[kH_feat + 6] = feat6;
        __syncwarp();
      }

      // ---- (i) SpeechNoiseProb (ns_core.c:642-749)
      {
        const float lrtAvg = NSB_FDIV_C(lsum, kMagnLenF);
        Hw[kH_feat + 3] = lrtAvg;
        // priorModelPars may have been re-estimated a few lines up (synced): read Hw
        const float thr0 = Hw[kH_priorPars + 0], thr1 = Hw[kH_priorPars + 1], thr2 = Hw[kH_priorPars + 3];
        const int sgn = (int)Hw[kH_priorPars + 2];
        // the three sigmoid maps take (float)tanh((double)x): one evaluation, lanes 0-2 an argument each
        float width = lrtAvg < thr0 ? 8.f : 4.f;
        const float a0 = width * (lrtAvg - thr0);
        const float sfv = feat0;
        width = 4.f;
        if (sgn == 1 && sfv > thr1) width = 8.f;
        if (sgn == -1 && sfv < thr1) width = 8.f;
        const float a1 = (float)sgn * width * (thr1 - sfv);
        const float sdv = feat4;
        width = sdv < thr2 ? 8.f : 4.f;
        const float a2 = width * (sdv - thr2);
        const float th = nsb_tanh_rn(lane == 0 ? a0 : (lane == 1 ? a1 : a2));
        const float ind0 = 0.5f * (__shfl_sync(kFullMask, th, 0) + 1.f);
        const float ind1 = 0.5f * (__shfl_sync(kFullMask, th, 1) + 1.f);
        const float ind2 = 0.5f * (__shfl_sync(kFullMask, th, 2) + 1.f);
        const float indPrior = Hw[kH_priorPars + 4] * ind0 + Hw[kH_priorPars + 5] * ind1 +
                               Hw[kH_priorPars + 6] * ind2;
        prior = Hr[kH_priorSpeechProb];
        prior += 0.1f * (indPrior - prior);
        if (prior > 1.f) prior = 1.f;
        if (prior < 0.01f) prior = 0.01f;
        Hw[kH_priorSpeechProb] = prior;
        const float gainPrior = fdiv(1.f - prior, prior + 0.0001f);
#pragma unroll
        for (int g = 0; g < G::kPairs; ++g) {
          const float2 pr = vfdiv(make_float2(1.f, 1.f), vmadd(make_float2(gainPrior, gainPrior), NSF_PAIR(invLrt, g), make_float2(1.f, 1.f)));
          NSF_UNPAIR(prob, g, pr);
        }
        prob[G::kSlots - 1] = fdiv(1.f, 1.f + gainPrior * invLrt[G::kSlots - 1]);
      }

      // ---- (j) UpdateNoiseEstimate (ns_core.c:800-846): gamma carried from bin i-1
      bool prevHigh[G::kSlots];   // the speech probability of bin k-1 exceeds 0.2 (false at bin 0)
#pragma unroll
      for (int j = 0; j < G::kSlots; ++j) {
        const bool nyq = (j == G::kSlots - 1);
        float pprev;
        if (nyq) {
          pprev = __shfl_sync(kFullMask, prob[G::kSlots - 2], 31);
        } else {
          pprev = __shfl_up_sync(kFullMask, prob[j], 1);
          const float wrap = __shfl_sync(kFullMask, prob[j > 0 ? j - 1 : 0], 31);
          if (lane == 0) pprev = wrap;
        }
        prevHigh[j] = !(!nyq && j == 0 && lane == 0) && pprev > 0.2f;
      }
#pragma unroll
      for (int g = 0; g < G::kPairs; ++g) {
        float2 mp = NSF_PAIR(mpause, g), nz;
        const bool2v ph = {prevHigh[2 * g], prevHigh[2 * g + 1]};
        nsf_noise_update(NSF_PAIR(prob, g), ph, NSF_PAIR(magn, g), NSF_PAIR(noisePrev, g), mp, nz);
        NSF_UNPAIR(mpause, g, mp);
        NSF_UNPAIR(noise, g, nz);
      }
      {
        constexpr int j = G::kSlots - 1;
        nsf_noise_update(prob[j], prevHigh[j], magn[j], noisePrev[j], mpause[j], noise[j]);
      }
      if (SPLIT) {
        // ns_core.c:1178-1180 and the fields Process reads later: noise, magnPrevAnalyze,
        // speechProb, parametricNoise; noisePrev stays as the last Process left it
#pragma unroll
        for (int j = 0; j < G::kSlots; ++j) {
          const bool nyq = (j == G::kSlots - 1);
          const int k = nyq ? G::kNC : lane + 32 * j;
          if (!nyq || lane == 0) {
            *reinterpret_cast<float4*>(B + k * kNsfBinRec + 8) = make_float4(noisePrev[j], magn[j], logLrt[j], mpause[j]);
            X_noise[k] = noise[j];
            X_prob[k] = prob[j];
            if (blockInd < 50) X_param[k] = parametric[j];
          }
        }
        __syncwarp();
      }
    }

    if (SPLIT && doP) {
      // ======== WebRtcNs_ProcessCore front end on its own signal (ns_core.c:1225-1267)
      blockInd = HIw[kH_blockInd];
      {
        float2 t[2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (lane + 32 * u < G::kHP) t[u] = blkP[G::kFP + lane + 32 * u];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (lane + 32 * u < G::kHP) blkP[lane + 32 * u] = t[u];
#pragma unroll
        for (int u = 0; u < kU; ++u)
          if (lane + 32 * u < G::kFP) blkP[G::kHP + lane + 32 * u] = pcm_unpack(cur[0][u]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = lane < G::kL ? blkP[lane + G::kL * j] : make_float2(0.f, 0.f);
      }
      energy1 = 0.f;
      if (lane < G::kL) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 w = reinterpret_cast<const float2*>(s_win)[lane + G::kL * j];
          v[j] = vmul(v[j], w);
          const float2 sq = vmul(v[j], v[j]);
          energy1 += sq.x + sq.y;
        }
      }
      energy1 = warp_sum(energy1);
      nzP = energy1 != 0.f;
      if (nzP) {
        ooura_fwd<G::kNC>(v, scr, s_otw, lane);
        if (lane < G::kL) {
#pragma unroll
          for (int q = 0; q < 4; ++q) scr[pad_idx(ooura_out_index<G::kNC>(lane, q))] = v[q];
        }
        __syncwarp();
        float re[G::kSlots], im[G::kSlots];
#pragma unroll
        for (int j = 0; j < G::kSlots; ++j) {
          const bool nyq = (j == G::kSlots - 1);
          const int k = nyq ? G::kNC : lane + 32 * j;
          const float2 zk = scr[pad_idx(k & (G::kNC - 1))];
          const float2 zm = scr[pad_idx((G::kNC - k) & (G::kNC - 1))];
          // the reference's real-input split (fft4g.c:1234-1256 and :352-354) seen from bin k: with
          // (wr, wi) = s_split[k] -- wi negated on the mirror side, (0, +-1/2) at bins 0 and N/2 so that
          // the same expression yields a[0] +- a[1] -- the same additions and products, rounding for rounding
          const float2 w = s_split[k];
          const float2 X = nsf_real_split(zk, zm, w);
          re[j] = X.x;
          im[j] = (nyq || k == 0) ? 0.f : X.y;
          magn[j] = nsb_sqrtf_p1(re[j] * re[j] + im[j] * im[j]);
          const float* R = B + k * kNsfBinRec;
          noise[j] = X_noise[k];
          prob[j] = X_prob[k];
          parametric[j] = X_param[k];
          // ComputeDdBasedWienerFilter's previous estimate (ns_core.c:993-994)
          prevEst[j] = fdiv(X_magnP[k], R[kB_noisePrev] + 0.0001f) * R[kB_smooth];
        }
        __syncwarp();  // scratch is free again
#pragma unroll
        for (int j = 0; j < G::kSlots; ++j) {
          const bool nyq = (j == G::kSlots - 1);
          if (!nyq || lane == 0) scr[nyq ? G::kNC : lane + 32 * j] = make_float2(re[j], im[j]);
        }
        prior = Hw[kH_priorSpeechProb];
      }
    }

    if (doP) {
    if (!nzP) {
      // ---- zero input to Process (ns_core.c:1239-1264): flush the overlap, high bands pass
      // through the delay line ungained.
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int pr = lane + 32 * u;
        o0[u] = make_float2(0.f, 0.f);
        if (pr < G::kHP) {   // a lane re-reads only what it wrote itself
          o0[u] = ovl[pr];
          ovl[pr] = make_float2(0.f, 0.f);
        }
      }
    } else {
      // ======== WebRtcNs_ProcessCore from the Wiener filter on
      // ---- (k) Wiener filter, flooring, start-up blend (ns_core.c:985-1007, 1268-1307)
      float gainW[G::kSlots];
#pragma unroll
      for (int g = 0; g < G::kPairs; ++g) {
        const float2 gw = nsf_wiener_gain(NSF_PAIR(magn, g), NSF_PAIR(noise, g), NSF_PAIR(prevEst, g), overdrive, denoiseBound);
        NSF_UNPAIR(gainW, g, gw);
      }
      gainW[G::kSlots - 1] = nsf_wiener_gain(magn[G::kSlots - 1], noise[G::kSlots - 1], prevEst[G::kSlots - 1], overdrive, denoiseBound);
      // the start-up blend sits outside the per-slot loops: a (warp-uniform) branch inside them
      // would fence the slots' division chains off from each other
      if (blockInd < 50) {
#pragma unroll
        for (int j = 0; j < G::kSlots; ++j) {
          const bool nyq = (j == G::kSlots - 1);
          const int k = nyq ? G::kNC : lane + 32 * j;
          float ime = gInitMagn[k] + magn[j];
          if (!nyq || lane == 0) gInitMagn[k] = ime;
          float ft = ime - overdrive * parametric[j];
          ft /= (ime + 0.0001f);
          if (ft < denoiseBound) ft = denoiseBound;
          if (ft > 1.f) ft = 1.f;
          float g = gainW[j] * (float)blockInd;
          ft *= (float)(50 - blockInd);
          g += ft;
          g /= 50.f;
          gainW[j] = g;
        }
      }
#pragma unroll
      for (int j = 0; j < G::kSlots; ++j) {
        const bool nyq = (j == G::kSlots - 1);
        const int k = nyq ? G::kNC : lane + 32 * j;
        const float flt = gainW[j];
        float2 z = scr[k];   // written by this lane (the Nyquist point by lane 0)
        z = vmul(z, make_float2(flt, flt));
        if (!nyq || lane == 0) {
          float* R = B + k * kNsfBinRec;
          R[kB_smooth] = flt;
          if (SPLIT) {
            // ns_core.c:1309-1310: magnPrevProcess = magn, noisePrev = noise
            if (NB > 1) {   // terms of sumMagnAnalyze / sumMagnProcess (:1372-1377), summed in bin order below
              stg[2 * G::kStg + k] = R[kB_magnPrev];
              stg[3 * G::kStg + k] = magn[j];
            }
            R[kB_noisePrev] = noise[j];
            X_magnP[k] = magn[j];
          } else {
            *reinterpret_cast<float4*>(R + 8) = make_float4(noise[j], magn[j], logLrt[j], mpause[j]);
          }
          scr[k] = z;
        }
        if (NB > 1) {
          // averages over the top quarter of the band, bins [magnLen - d - 1, magnLen - 1): one slot of all 32
          // lanes; its terms go to the staging arrays and are summed in bin order in (n)
          constexpr int d = G::kBins / 4;
          static_assert(NB == 1 || (d == 32 && (G::kBins - d - 1) % 32 == 0), "the high-band averages cover one slot");
          if (j == (G::kBins - d - 1) / 32) {
            stg[lane] = prob[j];
            stg[G::kStg + lane] = flt;
          }
        }
      }
      __syncwarp();

      // ---- (l) inverse FFT (ns_core.c:923-944)
      if (lane < G::kL) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = lane + G::kL * j;
          const float2 xk = scr[k];
          const float2 xm = scr[G::kNC - k];
          const float2 w = s_tw[k * (256 / ANA)];
          // (the four halves of this step and the 2 / N of the inverse transform are powers of two:
          // they commute with every rounding of the transform and are applied once, with the gain
          // factor, where the block is windowed -- kIfftScale below; same bits, 22 operations less)
          const float er = xk.x + xm.x, ei = xk.y - xm.y;
          const float dr = xk.x - xm.x, di = xk.y + xm.y;
          // O = (dr + i di) * conj(w);  Z = E + i O
          const float orr = dr * w.x + di * w.y, oi = di * w.x - dr * w.y;
          v[j] = make_float2(er - oi, ei + orr);
        }
      }
      __syncwarp();
      warp_fft<G::kNC, -1>(v, scr, s_tw, s_tw12, lane);
      if (lane < G::kL) {
#pragma unroll
        for (int q = 0; q < 4; ++q) scr[pad_idx(fft_out_index<G::kNC>(lane, q))] = v[q];
      }
      __syncwarp();

      // ---- (m) gain compensation, window, overlap-add (ns_core.c:1314-1359)
      constexpr int kTU = G::kTP / 32;  // pairs per lane over the whole block (4 / 2)
      float2 y[kTU];
      float energy2 = 0.f;
#pragma unroll
      for (int u = 0; u < kTU; ++u) {
        y[u] = scr[pad_idx(lane + 32 * u)];
        const float2 sq = vmul(y[u], y[u]);
        energy2 += sq.x + sq.y;
      }
      constexpr float kIfftScale = 1.f / (float)ANA;   // (2 / N) / 2: what y above still lacks
      float factor = 1.f;
      if (gainmap == 1 && blockInd > 200) {
        energy2 = warp_sum(energy2) * (kIfftScale * kIfftScale);
        float factor1 = 1.f, factor2 = 1.f;
        float gain = nsb_sqrtf(fdiv(energy2, energy1 + 1.f));   // branch-free forms: ns_warp.cuh
        if (gain > 0.5f) {
          factor1 = 1.f + 1.3f * (gain - 0.5f);
          if (gain * factor1 > 1.f) factor1 = 1.f / gain;
        }
        if (gain < 0.5f) {
          if (gain <= denoiseBound) gain = denoiseBound;
          factor2 = 1.f - 0.3f * (0.5f - gain);
        }
        factor = prior * factor1 + (1.f - prior) * factor2;
      }
      // windowed, scaled pairs back to scratch so that the tail can be re-read
      // in overlap order
      const float fscaled = factor * kIfftScale;
#pragma unroll
      for (int u = 0; u < kTU; ++u) {
        const float2 w = reinterpret_cast<const float2*>(s_win)[lane + 32 * u];
        y[u] = vmul(make_float2(fscaled, fscaled), vmul(w, y[u]));
      }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < kTU; ++u) scr[lane + 32 * u] = y[u];
      __syncwarp();
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int pr = lane + 32 * u;
        float2 o = make_float2(0.f, 0.f);
        if (pr < G::kFP) {
          o = scr[pr];
          if (pr < G::kHP) {
            const float2 t = ovl[pr];
            o = vadd(o, t);
            ovl[pr] = scr[G::kFP + pr];
          }
        }
        o0[u] = I16 ? o : make_float2(sat_s16f(o.x), sat_s16f(o.y));  // round_s16 saturates too
      }
      __syncwarp();

      // ---- (n) high-band time-domain gain (ns_core.c:1362-1404)
      if (NB > 1) {
        constexpr int d = G::kBins / 4;
        float hbProbSum, hbGainSum;
        {
          ChainSum4<32, G::kStg> ch(stg, lane);   // (staged in (k); a __syncwarp lies in between)
          const float c = ch.finish();
          hbProbSum = __shfl_sync(kFullMask, c, 0);
          hbGainSum = __shfl_sync(kFullMask, c, 1);
        }
        float avgProb = hbProbSum / (float)d;
        // sumMagnProcess / sumMagnAnalyze == 1 when Analyze and Process see one frame
        if (SPLIT) {
          ChainSum4<G::kStg, G::kStg> ch(stg, lane);
          const float c = ch.finish();
          avgProb *= __shfl_sync(kFullMask, c, 3) / __shfl_sync(kFullMask, c, 2);
        }
        const float avgGain = hbGainSum / (float)d;
        const float tmp = 2.f * avgProb - 1.f;
        const float gmod = 0.5f * (1.f + nsb_tanh_rn(tmp));
        float g = 0.5f * gmod + 0.5f * avgGain;
        if (avgProb >= 0.5f) g = 0.25f * gmod + 0.75f * avgGain;
        if (g < denoiseBound) g = denoiseBound;
        if (g > 1.f) g = 1.f;
        hbGain = g;
        hbApplyGain = true;
      }
    }

    // ---- outputs
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (lane + 32 * u < G::kFP) store_pair(f, 0, lane + 32 * u, o0[u]);

    if (NB > 1) {
      // delay line of 2*kHP samples per high band (dataBufHB, ns_core.c:1227-1235);
      // output = oldest kFrame samples of [history | new frame]
#pragma unroll
      for (int b = 0; b < NB - 1; ++b) {
        float2* blk = blkH + 128 * b;   // [delayed history | new frame]
        float2 t[2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (lane + 32 * u < G::kHP) t[u] = blk[G::kFP + lane + 32 * u];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (lane + 32 * u < G::kHP) blk[lane + 32 * u] = t[u];
#pragma unroll
        for (int u = 0; u < kU; ++u)
          if (lane + 32 * u < G::kFP) blk[G::kHP + lane + 32 * u] = pcm_unpack(cur[b + 1 < NB ? b + 1 : 0][u]);
        __syncwarp();
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int pr = lane + 32 * u;
          if (pr < G::kFP) {
            float2 o = blk[pr];
            if (hbApplyGain) o = vmul(o, make_float2(hbGain, hbGain));
            store_pair(f, b + 1, pr, I16 ? o : make_float2(sat_s16f(o.x), sat_s16f(o.y)));
          }
        }
      }
      __syncwarp();
    }
    }   // doP

#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int u = 0; u < kU; ++u) cur[b][u] = nxt[b][u];
    if (SPLIT) {
#pragma unroll
      for (int u = 0; u < kU; ++u) curA[u] = nxtA[u];
    }
    __syncwarp();
    { float* t = Hr; Hr = Hw; Hw = t; }
  }

  // ---- state: shared -> HBM, again as bulk copies issued by one lane
  if (!state_ready) mbar_wait_warp(barB, 0);
  // the fused kernel does not maintain the split-mode arrays: mark them stale
  if (lane == kH_splitValid && p.frames > 0) reinterpret_cast<int*>(Hr)[kH_splitValid] = SPLIT ? 1 : 0;
  bulk_store_fence();
  __syncwarp();
  if (lane == 0) {
    bulk_store(gS, Hr, kNsfHdrWords * 4);
    bulk_store(gS + kNsfOffXHist, blkA + G::kFP, kHistBytes);
    bulk_store(gS + kNsfOffSynt, ovl, kHistBytes);
    if (SPLIT) bulk_store(gS + kNsfOffPHist, blkP + G::kFP, kHistBytes);
#pragma unroll
    for (int b = 0; b < NB - 1; ++b) bulk_store(gS + kNsfOffHb + 96 * b, blkH + 128 * b + G::kFP, kHistBytes);
    bulk_store(gS + kNsfOffBins, B, kBinBytes);
    if (SPLIT) bulk_store(gS + kNsfOffAux, X_noise, 4 * kNsfAuxStride * 4);
    bulk_store_drain();
  }
}

}  // namespace nsb200

#endif  // AUDIOSIGNALPROCESS_B200_NSF_KERNEL_CUH_
