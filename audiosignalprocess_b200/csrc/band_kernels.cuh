// Band split / merge around the suppressors at 32 and 48 kHz, on the GPU:
//   32 kHz: one two-band QMF          (SplittingFilter::TwoBandsAnalysis/Synthesis,
//                                      modules/audio_processing/splitting_filter.cc:65-89)
//   48 kHz: sinc-resample 480 -> 640, three QMFs in a tree, top band dropped;
//           inverse on the way back  (ThreeBandsAnalysis/Synthesis, :95-171)
// replacing WebRtcSpl_AnalysisQMF / SynthesisQMF / AllPassQMF
// (common_audio/signal_processing/splitting_filter_c.c:48,127,167) and
// PushSincResampler / SincResampler::Resample / Convolve_SSE
// (common_audio/resampler/push_sinc_resampler.cc:33-100, sinc_resampler.cc:269-342,
// sinc_resampler_sse.cc:20-57).  Integer parts are bit-exact; the resampler
// follows the x86 reference's float operation order (4 interleaved partial
// sums, interpolate, (s0+s2)+(s1+s3)) so its int16 output is bit-exact too.
//
// Kernels (each walks all F frames of its unit, state in registers):
//   resample_*_kernel warp per stream, lanes over output samples
//   qmf_*_kernel      thread per (stream, QMF instance, all-pass chain): the
//                     cascaded all-pass sections are a non-linear serial
//                     recurrence over samples (saturating subtract, truncating
//                     32x16 multiply), so parallelism comes from streams x chains
#ifndef AUDIOSIGNALPROCESS_B200_BAND_KERNELS_CUH_
#define AUDIOSIGNALPROCESS_B200_BAND_KERNELS_CUH_

#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "band_layout.h"
#include "ns_warp.cuh"

namespace nsb200 {

NSB_DEV int band_sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
// the low (sel 0x9910) or high (sel 0xBB32) int16 of a word, sign-extended: one byte permute whose selector
// replicates the sign of the half's upper byte
NSB_DEV int band_half_s16(uint32_t w, uint32_t sel) {
#ifdef __CUDA_ARCH__
  int r;   // (prmt directly: __byte_perm does not promise the sign-replication bit of the selector)
  asm("prmt.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(w), "r"(sel));
  return r;
#else
  return sel == 0x9910u ? (int)(int16_t)(w & 0xffffu) : (int)(int16_t)(w >> 16);
#endif
}
NSB_DEV int band_round_s16(float v) {  // FloatS16ToS16, common_audio/include/audio_util.h:41-49
  if (v > 0.f) return v >= 32766.5f ? 32767 : (int)(v + 0.5f);
  return v <= -32767.5f ? -32768 : (int)(v - 0.5f);
}
// WEBRTC_SPL_SCALEDIFF32(A, B, C) = C + (B >> 16) * A + (((uint32_t)(B & 0xFFFF) * A) >> 16)
// (signal_processing_library.h:87-88).  With B = 65536 hi + lo the two products are hi*A and
// floor(lo*A / 65536), neither of which overflows for A < 2^16, so their sum is exactly
// floor(A*B / 65536): one wide multiply and a funnel shift on the recurrence's critical path.
NSB_DEV int band_scalediff(int a, int b, int c) {
#ifdef __CUDA_ARCH__
  long long prod;
  asm("mul.wide.s32 %0, %1, %2;" : "=l"(prod) : "r"(b), "r"(a));
#else
  const long long prod = (long long)b * (long long)a;
#endif
  return (int)((unsigned)c + (unsigned)(prod >> 16));
}

// One sample through the three cascaded first-order all-pass sections of
// WebRtcSpl_AllPassQMF (splitting_filter_c.c:48-105).  st[0..5] as in the
// reference: x[-1], y1[-1], y1[-1], y2[-1], y2[-1], y3[-1].
//
// The reference forms x - y[-1] with WebRtcSpl_SubSatW32 (spl_inl.h:60).  That saturation can
// never act: a section y[n] = x[n-1] + c (x[n] - y[n-1]) has the impulse response
// {c, 1-c^2, -c(1-c^2), ...} with absolute sum 1 + 2c, so from zero state |y3| <= 12.05 |x|max
// through the larger coefficient set (1.651 * 2.497 * 2.923) and 6.99 through the other.  Inputs
// are int16 sums in Q10, |x| <= 2^26 (synthesis: low +- high), hence every state stays below
// 8.1e8 and every difference below 1.7e9 < 2^31 (truncation adds < 1 per step).  A plain
// subtraction is therefore bit-identical and takes four instructions off the serial chain.
NSB_DEV int band_allpass3(int x, int (&st)[6], int c0, int c1, int c2) {
  const int y1 = band_scalediff(c0, x - st[1], st[0]);
  st[0] = x;
  st[1] = y1;
  const int y2 = band_scalediff(c1, y1 - st[3], st[2]);
  st[2] = y1;
  st[3] = y2;
  const int y3 = band_scalediff(c2, y2 - st[5], st[4]);
  st[4] = y2;
  st[5] = y3;
  return y3;
}

// ---- analysis QMF -------------------------------------------------------------
// Unit u = (stream, instance): input of `len` samples per frame, outputs low
// and high halves.  Two adjacent lanes are the two all-pass chains of one unit:
// chain 0 = odd samples through kAllPassFilter1, chain 1 = even samples through
// kAllPassFilter2 (splitting_filter_c.c:142-154); they swap results by shuffle.
struct QmfAnaLaunch {
  int32_t* state;       // band state slab
  const int* slots;
  int n_streams, frames, instances, len;   // len = input samples per frame per instance
  int state_off[2];     // word offset of analysis_state1 of each instance (state2 = +6)
  const int16_t* in[2]; // per instance
  long long in_stream_stride[2], in_frame_stride[2];
  int16_t* low[2];      // may be NULL (output discarded)
  int16_t* high[2];
  long long low_stream_stride[2], low_frame_stride[2], high_stream_stride[2], high_frame_stride[2];
};

constexpr int kQmfDepth = 4;   // prefetch ring depth (groups of 8 samples); divides 20 and 40

__global__ void __launch_bounds__(128) qmf_analysis_kernel(const QmfAnaLaunch p) {
  const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int chain = t & 1;
  const int unit = t >> 1;
  const bool live = unit < p.n_streams * p.instances;
  const int inst = live ? unit % p.instances : 0;
  const int sidx = live ? unit / p.instances : 0;
  const int c0 = chain == 0 ? 6418 : 21333, c1 = chain == 0 ? 36982 : 49062, c2 = chain == 0 ? 57261 : 63010;
  // per-thread constants that replace selects in the sample loop: which int16 of a word this chain filters
  // (byte selector with sign replication), and the sign of its own output in low = f1 + f2, high = f1 - f2
  const uint32_t half_sel = chain == 0 ? 0xBB32u : 0x9910u;
  const int own_sign = chain == 0 ? 1 : -1;
  int st[6] = {0, 0, 0, 0, 0, 0};
  int32_t* gst = nullptr;
  if (live) {
    gst = p.state + (size_t)p.slots[sidx] * kBandStateWords + p.state_off[inst] + 6 * chain;
#pragma unroll
    for (int i = 0; i < 6; ++i) st[i] = gst[i];
  }
  // Groups of 8 input pairs (32 bytes, read by both chains of the unit) are fetched kQmfDepth
  // groups ahead into a register ring: the recurrence spends ~200 cycles on a group, a global
  // load takes several times that.
  const int gpf = p.len / 16;                 // groups per frame (40 / 20)
  const int total = p.frames * gpf;           // multiple of kQmfDepth
  const int16_t* in_base = p.in[inst] + (size_t)sidx * p.in_stream_stride[inst];
  int16_t* lo_base = p.low[inst] ? p.low[inst] + (size_t)sidx * p.low_stream_stride[inst] : nullptr;
  int16_t* hi_base = p.high[inst] ? p.high[inst] + (size_t)sidx * p.high_stream_stride[inst] : nullptr;
  const long long in_fs = p.in_frame_stride[inst];
  const long long out_fs = chain == 0 ? p.low_frame_stride[inst] : p.high_frame_stride[inst];
  int16_t* dst_base = chain == 0 ? lo_base : hi_base;
  uint4 qa[kQmfDepth], qb[kQmfDepth];
  // fetch / store cursors advance group by group (no division): position in the frame, frame base
  int fetch_i = 0, store_i = 0;
  const int16_t* fetch_row = in_base;
  int16_t* store_row = dst_base;
  auto fetch = [&](int g, uint4& a, uint4& b) {
    a = b = make_uint4(0u, 0u, 0u, 0u);
    if (live && g < total) {
      const uint4* src = reinterpret_cast<const uint4*>(fetch_row + 16 * fetch_i);
      a = src[0];
      b = src[1];
    }
    if (++fetch_i == gpf) { fetch_i = 0; fetch_row += in_fs; }
  };
#pragma unroll
  for (int d = 0; d < kQmfDepth; ++d) fetch(d, qa[d], qb[d]);
  for (int g0 = 0; g0 < total; g0 += kQmfDepth) {
#pragma unroll
    for (int d = 0; d < kQmfDepth; ++d) {
      const int g = g0 + d;
      const uint4 a = qa[d], b = qb[d];
      fetch(g + kQmfDepth, qa[d], qb[d]);
      const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      uint32_t packed[4];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        // chain 0 filters in[2i+1], chain 1 filters in[2i]; both in Q10
        const int x = band_half_s16(w[k], half_sel) * 1024;
        const int y = band_allpass3(x, st, c0, c1, c2);
        const int other = __shfl_xor_sync(kFullMask, y, 1);
        // chain 0 emits the low band (f1 + f2 = y + other), chain 1 the high band (f1 - f2 = other - y) (:156-163)
        const int v = band_sat16((other + (y * own_sign + 1024)) >> 11);
        if (k & 1) packed[k >> 1] |= (uint32_t)v << 16;
        else packed[k >> 1] = (uint32_t)v & 0xffffu;
      }
      if (live && dst_base)
        *reinterpret_cast<uint4*>(store_row + 8 * store_i) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      if (++store_i == gpf) { store_i = 0; store_row += out_fs; }
    }
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < 6; ++i) gst[i] = st[i];
  }
}

// ---- synthesis QMF ------------------------------------------------------------
// out[2i] from chain 1 (kAllPassFilter1 on low-high), out[2i+1] from chain 0
// (kAllPassFilter2 on low+high) (splitting_filter_c.c:167-208).
struct QmfSynLaunch {
  int32_t* state;
  const int* slots;
  int n_streams, frames, instances, band_len;
  int state_off[2];      // word offset of synthesis_state1 (state2 = +6)
  const int16_t* low[2]; // NULL = all-zero band
  const int16_t* high[2];
  long long low_stream_stride[2], low_frame_stride[2], high_stream_stride[2], high_frame_stride[2];
  int16_t* out[2];
  long long out_stream_stride[2], out_frame_stride[2];
};

__global__ void __launch_bounds__(128) qmf_synthesis_kernel(const QmfSynLaunch p) {
  const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int chain = t & 1;
  const int unit = t >> 1;
  const bool live = unit < p.n_streams * p.instances;
  const int inst = live ? unit % p.instances : 0;
  const int sidx = live ? unit / p.instances : 0;
  // chain 0: state1 with kAllPassFilter2 on (low+high); chain 1: state2 with kAllPassFilter1 on (low-high)
  const int c0 = chain == 0 ? 21333 : 6418, c1 = chain == 0 ? 49062 : 36982, c2 = chain == 0 ? 63010 : 57261;
  const int high_sign = chain == 0 ? 1 : -1;   // input low + high (chain 0) or low - high (chain 1)
  const uint32_t pair_sel = chain == 0 ? 0x1054u : 0x5410u;   // byte selector of (even, odd) from (own, other)
  int st[6] = {0, 0, 0, 0, 0, 0};
  int32_t* gst = nullptr;
  if (live) {
    gst = p.state + (size_t)p.slots[sidx] * kBandStateWords + p.state_off[inst] + 6 * chain;
#pragma unroll
    for (int i = 0; i < 6; ++i) st[i] = gst[i];
  }
  const int gpf = p.band_len / 8;             // groups of 8 low + 8 high samples per frame (20 / 40)
  const int total = p.frames * gpf;           // multiple of kQmfDepth
  const int16_t* lo_base = (live && p.low[inst]) ? p.low[inst] + (size_t)sidx * p.low_stream_stride[inst] : nullptr;
  const int16_t* hi_base = (live && p.high[inst]) ? p.high[inst] + (size_t)sidx * p.high_stream_stride[inst] : nullptr;
  int16_t* out_base = live ? p.out[inst] + (size_t)sidx * p.out_stream_stride[inst] : nullptr;
  const long long lo_fs = p.low_frame_stride[inst], hi_fs = p.high_frame_stride[inst], out_fs = p.out_frame_stride[inst];
  uint4 ql[kQmfDepth], qh[kQmfDepth];
  int fetch_i = 0, store_i = 0;
  const int16_t* lo_row = lo_base;
  const int16_t* hi_row = hi_base;
  int16_t* out_row = out_base;
  auto fetch = [&](int g, uint4& l, uint4& h) {
    l = h = make_uint4(0u, 0u, 0u, 0u);
    if (g < total) {
      if (lo_base) l = *reinterpret_cast<const uint4*>(lo_row + 8 * fetch_i);
      if (hi_base) h = *reinterpret_cast<const uint4*>(hi_row + 8 * fetch_i);
    }
    if (++fetch_i == gpf) { fetch_i = 0; lo_row += lo_fs; hi_row += hi_fs; }
  };
#pragma unroll
  for (int d = 0; d < kQmfDepth; ++d) fetch(d, ql[d], qh[d]);
  for (int g0 = 0; g0 < total; g0 += kQmfDepth) {
#pragma unroll
    for (int d = 0; d < kQmfDepth; ++d) {
      const int g = g0 + d;
      const uint4 lw = ql[d], hw = qh[d];
      fetch(g + kQmfDepth, ql[d], qh[d]);
      const uint32_t l[4] = {lw.x, lw.y, lw.z, lw.w}, h[4] = {hw.x, hw.y, hw.z, hw.w};
      uint32_t ow[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int lv = (int)(int16_t)((k & 1) ? (l[k >> 1] >> 16) : (l[k >> 1] & 0xffffu));
        const int hv = (int)(int16_t)((k & 1) ? (h[k >> 1] >> 16) : (h[k >> 1] & 0xffffu));
        const int x = (hv * high_sign + lv) * 1024;
        const int y = band_allpass3(x, st, c0, c1, c2);
        const int v = band_sat16((y + 512) >> 10);
        const int other = __shfl_xor_sync(kFullMask, v, 1);
        // out[2i] = chain 1 (filter2), out[2i+1] = chain 0 (filter1): chain 0 packs (other, own), chain 1 (own, other)
#ifdef __CUDA_ARCH__
        ow[k] = __byte_perm((uint32_t)v, (uint32_t)other, pair_sel);
#else
        const int even = chain == 0 ? other : v, odd = chain == 0 ? v : other;
        ow[k] = ((uint32_t)even & 0xffffu) | ((uint32_t)odd << 16);
#endif
      }
      // the pair of lanes holds the same 8 words: chain 0 stores the first half
      if (out_base) {
        int16_t* out = out_row + 16 * store_i;
        if (chain == 0) *reinterpret_cast<uint4*>(out) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        else *reinterpret_cast<uint4*>(out + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
      }
      if (++store_i == gpf) { store_i = 0; out_row += out_fs; }
    }
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < 6; ++i) gst[i] = st[i];
  }
}

// ---- sinc resampler -------------------------------------------------------------
// One warp per stream.  E = [64 history samples | SRC new samples] as floats in
// shared memory; output n is a 32-tap dot product at position epos in E with
// kernels k1/k2 = table rows `off`, `off`+1 and linear interpolation between them.
//   UP   (480 -> 640, ratio 0.75, exact in binary): e_n = 31.5 + 0.75 n, closed form,
//        interpolation factor 0.
//   DOWN (640 -> 480, ratio 4/3 as a double): the reference advances a running double by
//        repeated addition (sinc_resampler.cc:322), so positions drift by rounding and depend
//        on the stream's age.  The host owns that double per handle, replays the reference's
//        block loop for the frames of this launch (band_host_init.h) and passes the resulting
//        schedule: per output (epos << 8 | off), (float)factor, (float)(1 - factor).
struct ResampleLaunch {
  int32_t* state;
  const int* slots;
  const float* kernel;        // 33 x 32 taps for this ratio
  const int32_t* schedule;    // DOWN only: [frames][480][3] words
  const int* stream_index;    // optional: launch-local index -> batch index (NULL = identity)
  const int16_t* in;
  int16_t* out;
  long long in_stream_stride, in_frame_stride, out_stream_stride, out_frame_stride;
  int n_streams, frames;
};

constexpr int kResampleWarpsPerCta = 4;
constexpr int kResampleKernelWords = 33 * 33 + 3;        // padded rows, 16-byte multiple
constexpr int kResampleEWords = 64 + 640 + 32;
constexpr size_t kResampleSmemBytes =
    sizeof(float) * (kResampleKernelWords + kResampleWarpsPerCta * kResampleEWords);
constexpr size_t kResampleUpSmemBytes = sizeof(float) * (kResampleWarpsPerCta * (64 + 480));

// 480 -> 640: positions are e_n = 31.5 + 0.75 n, so only table rows 0, 8, 16, 24
// are ever used and the interpolation factor is exactly 0.  The four rows live in constant
// memory: every tap is an FMUL with a constant-bank operand, no load at all.
__constant__ float c_up_rows[4][32];
// the same rows as pairs of adjacent taps: [0][row][k] = (c[2k], c[2k+1]), [1][row][k] = (c[2k+1], c[2k+2]) (k < 15)
__constant__ float2 c_up_pairs[2][4][16];

// Lane owns the 20 consecutive outputs n = 20 lane + t: their windows start at
// E[15 lane + ((126 + 3t) >> 2)], so 46 input samples held in registers serve all 20 outputs
// (46 conflict-free LDS per lane and frame instead of 640).  The next frame's PCM is fetched
// into registers while the current one is filtered.
__global__ void __launch_bounds__(kResampleWarpsPerCta * 32)
resample_up_kernel(const ResampleLaunch p) {
  extern __shared__ float4 rs_smem4[];
  float* smem = reinterpret_cast<float*>(rs_smem4);
  const int lane = lane_id(), warp = (int)(threadIdx.x >> 5);
  const int sidx = (int)blockIdx.x * kResampleWarpsPerCta + warp;
  if (sidx >= p.n_streams) return;
  float* E = smem + warp * (64 + 480);
  int32_t* gst = p.state + (size_t)p.slots[sidx] * kBandStateWords;
  int16_t* ghist = reinterpret_cast<int16_t*>(gst + kBandOffAnaHist);
  for (int i = lane; i < 64; i += 32) E[i] = (float)ghist[i];
  const int16_t* in = p.in + (size_t)sidx * p.in_stream_stride;
  int16_t* out = p.out + (size_t)sidx * p.out_stream_stride;
  constexpr int kW = 8;   // 240 words per frame over 32 lanes
  uint32_t nxt[kW];
  auto fetch = [&](int f) {
#pragma unroll
    for (int u = 0; u < kW; ++u) {
      const int w = lane + 32 * u;
      nxt[u] = (f < p.frames && w < 240) ? reinterpret_cast<const uint32_t*>(in + (size_t)f * p.in_frame_stride)[w] : 0u;
    }
  };
  fetch(0);
  for (int f = 0; f < p.frames; ++f) {
#pragma unroll
    for (int u = 0; u < kW; ++u) {
      const int w = lane + 32 * u;
      if (w < 240)
        *reinterpret_cast<float2*>(E + 64 + 2 * w) =
            make_float2((float)(int16_t)(nxt[u] & 0xffffu), (float)(int16_t)(nxt[u] >> 16));
    }
    fetch(f + 1);
    __syncwarp();
    float x[46];
#pragma unroll
    for (int j = 0; j < 46; ++j) x[j] = E[15 * lane + 31 + j];
    uint32_t packed[10] = {};
    // The four SSE-lane sums a0..a3 (taps i = q mod 4) advance two at a time as packed pairs of ADJACENT taps:
    // where the window starts at an even register the pairs are (a0, a1) and (a2, a3); where it starts at an
    // odd one they are (a1, a2) and (a3 at tap i, a0 at tap i + 4), with a0's first and a3's last term scalar
    // -- either way every sum sees its terms in the reference's order.  Tap pairs come from the constant
    // bank (c_up_pairs: both alignments); outputs are visited grouped by (table row, alignment) so that a pair
    // fetched once serves the two or three outputs that use it.
#pragma unroll
    for (int grp = 0; grp < 8; ++grp) {
#pragma unroll
      for (int t = 0; t < 20; ++t) {
        const int rel = ((126 + 3 * t) >> 2) - 31;        // window start relative to x[0]
        const int row = (126 + 3 * t) & 3;                // table row 8 * row = c_up_rows[row]
        if (((rel & 1) * 4 + row) != grp) continue;
        float a0, a1, a2, a3;
        if ((rel & 1) == 0) {
          float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            s01 = vadd(s01, vmul_o(make_float2(x[rel + i], x[rel + i + 1]), c_up_pairs[0][row][i / 2]));
            s23 = vadd(s23, vmul_o(make_float2(x[rel + i + 2], x[rel + i + 3]), c_up_pairs[0][row][i / 2 + 1]));
          }
          a0 = s01.x; a1 = s01.y; a2 = s23.x; a3 = s23.y;
        } else {
          float2 s12 = make_float2(0.f, 0.f), s30 = make_float2(0.f, 0.f + x[rel] * c_up_rows[row][0]);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            s12 = vadd(s12, vmul_o(make_float2(x[rel + i + 1], x[rel + i + 2]), c_up_pairs[1][row][i / 2]));
            if (i < 28) s30 = vadd(s30, vmul_o(make_float2(x[rel + i + 3], x[rel + i + 4]), c_up_pairs[1][row][i / 2 + 1]));
          }
          a0 = s30.y; a1 = s12.x; a2 = s12.y;
          a3 = s30.x + x[rel + 31] * c_up_rows[row][31];
        }
        // sinc_resampler_sse.cc:42-54 with interpolation factor 0: a*1 + b*0 == a for the finite
        // sums b of the second kernel (up to the sign of zero, gone after rounding)
        const int v = band_round_s16((a0 + a2) + (a1 + a3));
        if (t & 1) packed[t >> 1] = (packed[t >> 1] & 0xffffu) | ((uint32_t)v << 16);
        else packed[t >> 1] = (packed[t >> 1] & 0xffff0000u) | ((uint32_t)v & 0xffffu);
      }
    }
    uint2* dst = reinterpret_cast<uint2*>(out + (size_t)f * p.out_frame_stride + 20 * lane);
#pragma unroll
    for (int q = 0; q < 5; ++q) dst[q] = make_uint2(packed[2 * q], packed[2 * q + 1]);
    // history: last 64 input samples
    const float h0 = E[480 + lane], h1 = E[480 + 32 + lane];
    __syncwarp();
    E[lane] = h0;
    E[32 + lane] = h1;
    __syncwarp();
  }
  for (int i = lane; i < 64; i += 32) ghist[i] = (int16_t)E[i];
}

// 640 -> 480: warp per stream, lanes over outputs n = lane + 32 t; the schedule and the PCM of the
// next frame are fetched into registers while the current frame is filtered.
__global__ void __launch_bounds__(kResampleWarpsPerCta * 32)
resample_down_kernel(const ResampleLaunch p) {
  constexpr int SRC = 640, DST = 480, NOUT = DST / 32;   // 15 outputs per lane
  extern __shared__ float4 rs_smem4[];
  float* smem = reinterpret_cast<float*>(rs_smem4);
  float* s_kernel = smem;                                  // 33 rows of 32 taps, row stride 33 (banks)
  const int lane = lane_id(), warp = (int)(threadIdx.x >> 5);
  for (int i = (int)threadIdx.x; i < 33 * 32; i += kResampleWarpsPerCta * 32)
    s_kernel[(i >> 5) * 33 + (i & 31)] = p.kernel[i];
  __syncthreads();
  const int lidx = (int)blockIdx.x * kResampleWarpsPerCta + warp;
  if (lidx >= p.n_streams) return;
  const int sidx = p.stream_index ? p.stream_index[lidx] : lidx;
  float* E = smem + kResampleKernelWords + warp * kResampleEWords;

  int32_t* gst = p.state + (size_t)p.slots[sidx] * kBandStateWords;
  int16_t* ghist = reinterpret_cast<int16_t*>(gst + kBandOffSynHist);
  for (int i = lane; i < 64; i += 32) E[i] = (float)ghist[i];
  const int16_t* in = p.in + (size_t)sidx * p.in_stream_stride;
  int16_t* outp = p.out + (size_t)sidx * p.out_stream_stride;
  constexpr int kW = SRC / 64;   // 10 words per lane
  uint32_t nxt[kW];
  int32_t sch[NOUT][3];
  auto fetch = [&](int f) {
    if (f < p.frames) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (size_t)f * p.in_frame_stride);
#pragma unroll
      for (int u = 0; u < kW; ++u) nxt[u] = src[lane + 32 * u];
      const int32_t* sp = p.schedule + (size_t)f * DST * 3;
#pragma unroll
      for (int t = 0; t < NOUT; ++t) {
        const int n = lane + 32 * t;
        sch[t][0] = sp[3 * n];
        sch[t][1] = sp[3 * n + 1];
        sch[t][2] = sp[3 * n + 2];
      }
    }
  };
  fetch(0);
  for (int f = 0; f < p.frames; ++f) {
    int16_t* out = outp + (size_t)f * p.out_frame_stride;
#pragma unroll
    for (int u = 0; u < kW; ++u) {
      const int w = lane + 32 * u;
      *reinterpret_cast<float2*>(E + 64 + 2 * w) =
          make_float2((float)(int16_t)(nxt[u] & 0xffffu), (float)(int16_t)(nxt[u] >> 16));
    }
    int32_t cur[NOUT][3];
#pragma unroll
    for (int t = 0; t < NOUT; ++t) { cur[t][0] = sch[t][0]; cur[t][1] = sch[t][1]; cur[t][2] = sch[t][2]; }
    fetch(f + 1);
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NOUT; ++t) {
      const int n = lane + 32 * t;
      const int epos = cur[t][0] >> 8, off = cur[t][0] & 0xff;
      const float fac = __int_as_float(cur[t][1]), fac1 = __int_as_float(cur[t][2]);
      const float* k1 = s_kernel + off * 33;
      const float* k2 = k1 + 33;
      const float* x = E + epos;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        a0 += x[i] * k1[i];
        a1 += x[i + 1] * k1[i + 1];
        a2 += x[i + 2] * k1[i + 2];
        a3 += x[i + 3] * k1[i + 3];
        b0 += x[i] * k2[i];
        b1 += x[i + 1] * k2[i + 1];
        b2 += x[i + 2] * k2[i + 2];
        b3 += x[i + 3] * k2[i + 3];
      }
      // sinc_resampler_sse.cc:42-54
      const float t0 = a0 * fac1 + b0 * fac, t1 = a1 * fac1 + b1 * fac;
      const float t2 = a2 * fac1 + b2 * fac, t3 = a3 * fac1 + b3 * fac;
      out[n] = (int16_t)band_round_s16((t0 + t2) + (t1 + t3));
    }
    __syncwarp();
    // history: last 64 input samples
    const float h0 = E[SRC + lane], h1 = E[SRC + 32 + lane];
    __syncwarp();
    E[lane] = h0;
    E[32 + lane] = h1;
    __syncwarp();
  }
  for (int i = lane; i < 64; i += 32) ghist[i] = (int16_t)E[i];
}

// 640 -> 480 on a REGULAR schedule.  The ratio is 4/3: output n starts its window at E[32 + floor(4 n / 3)]
// and uses table rows off(n mod 3) = 0, 10, 21 (and the row after) for as long as the reference's running
// double stays within 1/96 sample of the exact position -- from a stream's second block on and for years
// (the double of 4/3 is rounded up: the drift is +8e-11 samples per frame; band_host_init.h replays it and
// the host checks every launch's schedule, so the general kernel above takes whatever is not regular: the
// priming frames of a new stream).  Regular positions turn the general kernel's 7 instructions per tap
// (three shared-memory loads, two products, two sums) into 2:
//   * lane L owns the 15 consecutive outputs n = 15 L + t.  Their windows lie inside the 50 samples
//     E[32 + 20 L ...], loaded once per frame into registers (13 conflict-free 16-byte loads);
//   * n mod 3 = t mod 3, so at every unrolled tap all lanes use the same two table rows: the (k1, k2) pairs
//     sit in the constant bank and are an operand of the multiply, no load at all;
//   * the sums over k1 and k2 advance together as one packed pair: FFMA2 (product, opaque to contraction:
//     ns_warp.cuh vmul_o) + FADD2 per tap for both.
// Every accumulator sees the reference's operations in the reference's order (sinc_resampler_sse.cc:20-57),
// so the int16 output is bit-identical to the general kernel's.
__constant__ float2 c_down_pairs[3][32];   // [n mod 3][tap] = (row off, row off + 1), off = 0, 10, 21
constexpr int kDownRegularRows[3] = {0, 10, 21};
constexpr int kDownRegularEpos = 32;
constexpr size_t kResampleDownRegSmemBytes = sizeof(float) * kResampleWarpsPerCta * (64 + 640 + 240);

#ifndef NSB_DOWN_MIN_CTAS
#define NSB_DOWN_MIN_CTAS 3   // 168 registers: the constant-bank tap pairs stay in registers across a phase (4: 127 registers, +170 MOV/LDC per frame, 3 % slower step)
#endif
__global__ void __launch_bounds__(kResampleWarpsPerCta * 32, NSB_DOWN_MIN_CTAS)
resample_down_regular_kernel(const ResampleLaunch p) {
  constexpr int SRC = 640, DST = 480, NOUT = 15;
  extern __shared__ float4 rs_smem4[];
  float* smem = reinterpret_cast<float*>(rs_smem4);
  const int lane = lane_id(), warp = (int)(threadIdx.x >> 5);
  const int lidx = (int)blockIdx.x * kResampleWarpsPerCta + warp;
  if (lidx >= p.n_streams) return;
  const int sidx = p.stream_index ? p.stream_index[lidx] : lidx;
  float* E = smem + warp * (64 + SRC + 240);
  uint32_t* O = reinterpret_cast<uint32_t*>(E + 64 + SRC);   // the frame's 480 outputs, for coalesced stores

  int32_t* gst = p.state + (size_t)p.slots[sidx] * kBandStateWords;
  int16_t* ghist = reinterpret_cast<int16_t*>(gst + kBandOffSynHist);
  for (int i = lane; i < 64; i += 32) E[i] = (float)ghist[i];
  const int16_t* in = p.in + (size_t)sidx * p.in_stream_stride;
  int16_t* outp = p.out + (size_t)sidx * p.out_stream_stride;
  constexpr int kW = SRC / 64;   // 10 words per lane
  uint32_t nxt[kW];
  auto fetch = [&](int f) {
    if (f < p.frames) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (size_t)f * p.in_frame_stride);
#pragma unroll
      for (int u = 0; u < kW; ++u) nxt[u] = src[lane + 32 * u];
    }
  };
  fetch(0);
  for (int f = 0; f < p.frames; ++f) {
#pragma unroll
    for (int u = 0; u < kW; ++u) {
      const int wd = lane + 32 * u;
      *reinterpret_cast<float2*>(E + 64 + 2 * wd) =
          make_float2((float)(int16_t)(nxt[u] & 0xffffu), (float)(int16_t)(nxt[u] >> 16));
    }
    const int32_t* sp = p.schedule + ((size_t)f * DST + NOUT * lane) * 3;
    fetch(f + 1);
    __syncwarp();
    float x[52];
#pragma unroll
    for (int j = 0; j < 13; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(E + kDownRegularEpos + 20 * lane + 4 * j);
      x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
    }
    // phase by phase: the five outputs t = ph, ph + 3, ... share their table rows, so each (k1, k2) pair
    // is fetched from the constant bank once per frame
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
      float2 acc[5][4], w[5];   // w: (1 - factor, factor) of the phase's outputs (the schedule is shared by the group: L1 hits)
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        w[u] = make_float2(__int_as_float(sp[3 * (ph + 3 * u) + 2]), __int_as_float(sp[3 * (ph + 3 * u) + 1]));
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[u][q] = make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 k12 = c_down_pairs[ph][i + q];
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int rel = (4 * (ph + 3 * u)) / 3;
            acc[u][q] = vadd(acc[u][q], vmul_o(make_float2(x[rel + i + q], x[rel + i + q]), k12));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int t = ph + 3 * u;
        // sinc_resampler_sse.cc:42-54: sums1 * (1 - factor) + sums2 * factor per SSE lane, then (s0 + s2) + (s1 + s3)
        float tq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 pr = vmul(acc[u][q], w[u]);
          tq[q] = pr.x + pr.y;
        }
        const uint32_t v = (uint32_t)band_round_s16((tq[0] + tq[2]) + (tq[1] + tq[3])) & 0xffffu;
        // output n = 15 lane + t: int16 pairs straddle lanes when lane is odd, so go through 16-bit stores
        reinterpret_cast<uint16_t*>(O)[NOUT * lane + t] = (uint16_t)v;
      }
    }
    __syncwarp();
    {
      uint32_t* dst = reinterpret_cast<uint32_t*>(outp + (size_t)f * p.out_frame_stride);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int wd = lane + 32 * u;
        if (wd < DST / 2) dst[wd] = O[wd];
      }
    }
    // history: last 64 input samples
    const float h0 = E[SRC + lane], h1 = E[SRC + 32 + lane];
    __syncwarp();
    E[lane] = h0;
    E[32 + lane] = h1;
    __syncwarp();
  }
  for (int i = lane; i < 64; i += 32) ghist[i] = (int16_t)E[i];
}

// Is frame f of a 640 -> 480 schedule (480 x 3 words, band_host_init.h) regular in the above sense?
inline bool band_down_frame_regular(const int32_t* sched) {
  for (int n = 0; n < 480; ++n) {
    const int32_t w0 = sched[3 * n];
    if ((w0 & 0xff) != kDownRegularRows[n % 3] || (w0 >> 8) != kDownRegularEpos + (4 * n) / 3) return false;
  }
  return true;
}

// ---- launch helpers (host) -------------------------------------------------------
// The band path is a chain of stages around the suppressor, each a kernel that walks the frames
// of its streams serially (filter / resampler state in registers).  Stage s of frames [f0, f0+nf)
// needs stage s-1 of the same frames and stage s of the frames before; the host layer
// (ns_capi.cu, RunDevice) issues chunks of frames through one CUDA stream per stage so that the
// stages of neighbouring chunks overlap.
//   48 kHz: 0 resample 480->640 | 1 QMF analysis 64k->2x32k | 2 QMF analysis 2x32k->3 bands |
//           3 NS | 4 QMF synthesis bands->2x32k | 5 QMF synthesis ->64k | 6 resample 640->480
//   32 kHz: 0 QMF analysis | 1 NS | 2 QMF synthesis
struct BandLaunch {
  int32_t* state;
  const int* slots;
  const float* kernel_up;     // resampler tables (device)
  const float* kernel_down;
  // 48 kHz merge: the 640 -> 480 resampler runs once per group of streams that share a
  // position schedule (normally one group = the whole batch)
  // regular[f]: frame f of the schedule can take the fast kernel (band_down_frame_regular), host memory
  struct DownGroup { const int32_t* schedule; const int* stream_index; int count; const uint8_t* regular; };
  std::vector<DownGroup> down_groups;
  const int16_t* full_in;     // full-band PCM [stream][frame][fs/100]
  long long full_in_stride;
  int16_t* full_out;
  long long full_out_stride;
  int16_t* bands;             // [stream][frame][nb][160]
  long long bands_stride;
  int16_t* scratch;           // 48 kHz: [2][stream][frame][640]
  int n_streams, frames;      // frames = all frames of the call (scratch layout)
};

inline int BandStages(int nb) { return nb == 3 ? 7 : 3; }
inline int BandNsStage(int nb) { return nb == 3 ? 3 : 1; }

inline size_t BandScratchElems(int nb, int n_streams, int frames) {
  return nb == 3 ? (size_t)2 * n_streams * frames * 640 : 0;
}

// Launches band stage `stage` (not the NS stage) for frames [f0, f0 + nf).
inline int LaunchBandStage(int nb, const BandLaunch& b, int stage, int f0, int nf, cudaStream_t st,
                           uint64_t* launches) {
  const int n = b.n_streams;
  const long long bs = b.bands_stride, fstride = nb * 160;
  int16_t* bands = b.bands + (size_t)f0 * fstride;
  const int16_t* fin = b.full_in + (size_t)f0 * (nb == 3 ? 480 : 320);
  int16_t* fout = b.full_out + (size_t)f0 * (nb == 3 ? 480 : 320);
  const int qgrid1 = (2 * n + 127) / 128, qgrid2 = (4 * n + 127) / 128;
  if (nb == 2) {
    if (stage == 0) {
      QmfAnaLaunch q = {};
      q.state = b.state; q.slots = b.slots; q.n_streams = n; q.frames = nf; q.instances = 1; q.len = 320;
      q.state_off[0] = kBandOffQmf0;
      q.in[0] = fin; q.in_stream_stride[0] = b.full_in_stride; q.in_frame_stride[0] = 320;
      q.low[0] = bands; q.low_stream_stride[0] = bs; q.low_frame_stride[0] = fstride;
      q.high[0] = bands + 160; q.high_stream_stride[0] = bs; q.high_frame_stride[0] = fstride;
      qmf_analysis_kernel<<<qgrid1, 128, 0, st>>>(q);
    } else if (stage == 2) {
      QmfSynLaunch q = {};
      q.state = b.state; q.slots = b.slots; q.n_streams = n; q.frames = nf; q.instances = 1; q.band_len = 160;
      q.state_off[0] = kBandOffQmf0 + 12;
      q.low[0] = bands; q.high[0] = bands + 160;
      q.low_stream_stride[0] = q.high_stream_stride[0] = bs;
      q.low_frame_stride[0] = q.high_frame_stride[0] = fstride;
      q.out[0] = fout; q.out_stream_stride[0] = b.full_out_stride; q.out_frame_stride[0] = 320;
      qmf_synthesis_kernel<<<qgrid1, 128, 0, st>>>(q);
    } else {
      return -1;
    }
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  if (nb != 3) return -1;
  const long long ss = (long long)b.frames * 640;
  int16_t* s64 = b.scratch + (size_t)f0 * 640;                                   // 64 kHz signal
  int16_t* s32 = b.scratch + (size_t)n * b.frames * 640 + (size_t)f0 * 640;      // [low 320 | high 320]
  const int rgrid = (n + kResampleWarpsPerCta - 1) / kResampleWarpsPerCta;
  switch (stage) {
    case 0: {
      ResampleLaunch r = {};
      r.state = b.state; r.slots = b.slots; r.kernel = b.kernel_up; r.in = fin; r.out = s64;
      r.in_stream_stride = b.full_in_stride; r.in_frame_stride = 480; r.out_stream_stride = ss; r.out_frame_stride = 640;
      r.n_streams = n; r.frames = nf;
      resample_up_kernel<<<rgrid, kResampleWarpsPerCta * 32, kResampleUpSmemBytes, st>>>(r);
      ++*launches;
      break;
    }
    case 1: {
      QmfAnaLaunch q = {};
      q.state = b.state; q.slots = b.slots; q.n_streams = n; q.frames = nf; q.instances = 1; q.len = 640;
      q.state_off[0] = kBandOffQmf0;
      q.in[0] = s64; q.in_stream_stride[0] = ss; q.in_frame_stride[0] = 640;
      q.low[0] = s32; q.low_stream_stride[0] = ss; q.low_frame_stride[0] = 640;
      q.high[0] = s32 + 320; q.high_stream_stride[0] = ss; q.high_frame_stride[0] = 640;
      qmf_analysis_kernel<<<qgrid1, 128, 0, st>>>(q);
      ++*launches;
      break;
    }
    case 2: {
      QmfAnaLaunch q2 = {};
      q2.state = b.state; q2.slots = b.slots; q2.n_streams = n; q2.frames = nf; q2.instances = 2; q2.len = 320;
      q2.state_off[0] = kBandOffQmf1; q2.state_off[1] = kBandOffQmf2;
      q2.in[0] = s32; q2.in[1] = s32 + 320;
      for (int i = 0; i < 2; ++i) { q2.in_stream_stride[i] = ss; q2.in_frame_stride[i] = 640; }
      q2.low[0] = bands; q2.high[0] = bands + 160;           // 0-8 kHz, 8-16 kHz
      q2.low[1] = nullptr; q2.high[1] = bands + 320;         // (24-32 kHz dropped), 16-24 kHz
      for (int i = 0; i < 2; ++i) {
        q2.low_stream_stride[i] = q2.high_stream_stride[i] = bs;
        q2.low_frame_stride[i] = q2.high_frame_stride[i] = fstride;
      }
      qmf_analysis_kernel<<<qgrid2, 128, 0, st>>>(q2);
      ++*launches;
      break;
    }
    case 4: {
      QmfSynLaunch q2 = {};
      q2.state = b.state; q2.slots = b.slots; q2.n_streams = n; q2.frames = nf; q2.instances = 2; q2.band_len = 160;
      q2.state_off[0] = kBandOffQmf1 + 12; q2.state_off[1] = kBandOffQmf2 + 12;
      q2.low[0] = bands; q2.high[0] = bands + 160;
      q2.low[1] = nullptr; q2.high[1] = bands + 320;         // zero low half (splitting_filter.cc:143-145,152)
      for (int i = 0; i < 2; ++i) {
        q2.low_stream_stride[i] = q2.high_stream_stride[i] = bs;
        q2.low_frame_stride[i] = q2.high_frame_stride[i] = fstride;
        q2.out_stream_stride[i] = ss; q2.out_frame_stride[i] = 640;
      }
      q2.out[0] = s32; q2.out[1] = s32 + 320;
      qmf_synthesis_kernel<<<qgrid2, 128, 0, st>>>(q2);
      ++*launches;
      break;
    }
    case 5: {
      QmfSynLaunch q = {};
      q.state = b.state; q.slots = b.slots; q.n_streams = n; q.frames = nf; q.instances = 1; q.band_len = 320;
      q.state_off[0] = kBandOffQmf0 + 12;
      q.low[0] = s32; q.high[0] = s32 + 320;
      q.low_stream_stride[0] = q.high_stream_stride[0] = ss;
      q.low_frame_stride[0] = q.high_frame_stride[0] = 640;
      q.out[0] = s64; q.out_stream_stride[0] = ss; q.out_frame_stride[0] = 640;
      qmf_synthesis_kernel<<<qgrid1, 128, 0, st>>>(q);
      ++*launches;
      break;
    }
    case 6: {
      for (const BandLaunch::DownGroup& g : b.down_groups) {
        ResampleLaunch r = {};
        r.state = b.state; r.slots = b.slots; r.kernel = b.kernel_down;
        r.schedule = g.schedule + (size_t)f0 * 480 * 3;
        r.stream_index = g.stream_index; r.in = s64; r.out = fout;
        r.in_stream_stride = ss; r.in_frame_stride = 640; r.out_stream_stride = b.full_out_stride; r.out_frame_stride = 480;
        r.n_streams = g.count; r.frames = nf;
        // NSB200_BAND_GENERAL_DOWN=1 keeps the general kernel (test hook: both must give the same bits)
        static const bool force_general = getenv("NSB200_BAND_GENERAL_DOWN") != nullptr;
        bool regular = g.regular != nullptr && !force_general;
        for (int f = f0; regular && f < f0 + nf; ++f) regular = g.regular[f] != 0;
        const int grid = (g.count + kResampleWarpsPerCta - 1) / kResampleWarpsPerCta;
        if (regular)
          resample_down_regular_kernel<<<grid, kResampleWarpsPerCta * 32, kResampleDownRegSmemBytes, st>>>(r);
        else
          resample_down_kernel<<<grid, kResampleWarpsPerCta * 32, kResampleSmemBytes, st>>>(r);
        ++*launches;
      }
      break;
    }
    default:
      return -1;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace nsb200

#endif  // AUDIOSIGNALPROCESS_B200_BAND_KERNELS_CUH_
