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
//   resample_kernel   warp per stream, lanes over output samples
//   qmf_*_kernel      thread per (stream, QMF instance, all-pass chain): the
//                     cascaded all-pass sections are a non-linear serial
//                     recurrence over samples (saturating subtract, truncating
//                     32x16 multiply), so parallelism comes from streams x chains
#ifndef AUDIOSIGNALPROCESS_B200_BAND_KERNELS_CUH_
#define AUDIOSIGNALPROCESS_B200_BAND_KERNELS_CUH_

#include <stdint.h>

#include <vector>

#include "band_layout.h"
#include "ns_warp.cuh"

namespace nsb200 {

NSB_DEV int band_sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
NSB_DEV int band_round_s16(float v) {  // FloatS16ToS16, common_audio/include/audio_util.h:41-49
  if (v > 0.f) return v >= 32766.5f ? 32767 : (int)(v + 0.5f);
  return v <= -32767.5f ? -32768 : (int)(v - 0.5f);
}
// WebRtcSpl_SubSatW32 (spl_inl.h:60)
NSB_DEV int band_sub_sat(int a, int b) {
  const int d = (int)((unsigned)a - (unsigned)b);
  // overflow iff the operands differ in sign and the result's sign differs from a's
  const bool ovf = ((a ^ b) & (a ^ d)) < 0;
  return ovf ? (a < 0 ? (int)0x80000000 : 0x7fffffff) : d;
}
// WEBRTC_SPL_SCALEDIFF32(A, B, C) = C + (B >> 16) * A + (((uint32_t)(B & 0xFFFF) * A) >> 16)
NSB_DEV int band_scalediff(unsigned a, int b, int c) {
  return (int)((unsigned)c + (unsigned)(b >> 16) * a + ((((unsigned)b & 0xFFFFu) * a) >> 16));
}

// One sample through the three cascaded first-order all-pass sections of
// WebRtcSpl_AllPassQMF (splitting_filter_c.c:48-105).  st[0..5] as in the
// reference: x[-1], y1[-1], y1[-1], y2[-1], y2[-1], y3[-1].
NSB_DEV int band_allpass3(int x, int (&st)[6], unsigned c0, unsigned c1, unsigned c2) {
  const int y1 = band_scalediff(c0, band_sub_sat(x, st[1]), st[0]);
  st[0] = x;
  st[1] = y1;
  const int y2 = band_scalediff(c1, band_sub_sat(y1, st[3]), st[2]);
  st[2] = y1;
  st[3] = y2;
  const int y3 = band_scalediff(c2, band_sub_sat(y2, st[5]), st[4]);
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

__global__ void __launch_bounds__(128) qmf_analysis_kernel(const QmfAnaLaunch p) {
  const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int chain = t & 1;
  const int unit = t >> 1;
  const bool live = unit < p.n_streams * p.instances;
  const int inst = live ? unit % p.instances : 0;
  const int sidx = live ? unit / p.instances : 0;
  const unsigned c0 = chain == 0 ? 6418u : 21333u, c1 = chain == 0 ? 36982u : 49062u,
                 c2 = chain == 0 ? 57261u : 63010u;
  int st[6] = {0, 0, 0, 0, 0, 0};
  int32_t* gst = nullptr;
  if (live) {
    gst = p.state + (size_t)p.slots[sidx] * kBandStateWords + p.state_off[inst] + 6 * chain;
#pragma unroll
    for (int i = 0; i < 6; ++i) st[i] = gst[i];
  }
  const int half = p.len / 2;
  for (int f = 0; f < p.frames; ++f) {
    const int16_t* in = p.in[inst] + (size_t)sidx * p.in_stream_stride[inst] + (size_t)f * p.in_frame_stride[inst];
    int16_t* lo = p.low[inst] ? p.low[inst] + (size_t)sidx * p.low_stream_stride[inst] + (size_t)f * p.low_frame_stride[inst] : nullptr;
    int16_t* hi = p.high[inst] ? p.high[inst] + (size_t)sidx * p.high_stream_stride[inst] + (size_t)f * p.high_frame_stride[inst] : nullptr;
    for (int i0 = 0; i0 < half; i0 += 8) {
      // 8 input pairs = 32 bytes, read by both chains of the unit
      uint32_t w[8];
      if (live) {
        const uint4 a = reinterpret_cast<const uint4*>(in + 2 * i0)[0];
        const uint4 b = reinterpret_cast<const uint4*>(in + 2 * i0)[1];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = 0u;
      }
      uint32_t packed[4];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        // chain 0 filters in[2i+1], chain 1 filters in[2i]; both in Q10
        const int x = (chain == 0 ? (int)(int16_t)(w[k] >> 16) : (int)(int16_t)(w[k] & 0xffffu)) * 1024;
        const int y = band_allpass3(x, st, c0, c1, c2);
        const int other = __shfl_xor_sync(kFullMask, y, 1);
        const int f1 = chain == 0 ? y : other, f2 = chain == 0 ? other : y;
        // chain 0 emits the low band, chain 1 the high band (:156-163)
        const int v = chain == 0 ? band_sat16((f1 + f2 + 1024) >> 11) : band_sat16((f1 - f2 + 1024) >> 11);
        if (k & 1) packed[k >> 1] |= (uint32_t)v << 16;
        else packed[k >> 1] = (uint32_t)v & 0xffffu;
      }
      int16_t* dst = chain == 0 ? lo : hi;
      if (live && dst) *reinterpret_cast<uint4*>(dst + i0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
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
  const unsigned c0 = chain == 0 ? 21333u : 6418u, c1 = chain == 0 ? 49062u : 36982u,
                 c2 = chain == 0 ? 63010u : 57261u;
  int st[6] = {0, 0, 0, 0, 0, 0};
  int32_t* gst = nullptr;
  if (live) {
    gst = p.state + (size_t)p.slots[sidx] * kBandStateWords + p.state_off[inst] + 6 * chain;
#pragma unroll
    for (int i = 0; i < 6; ++i) st[i] = gst[i];
  }
  for (int f = 0; f < p.frames; ++f) {
    const int16_t* lo = (live && p.low[inst]) ? p.low[inst] + (size_t)sidx * p.low_stream_stride[inst] + (size_t)f * p.low_frame_stride[inst] : nullptr;
    const int16_t* hi = (live && p.high[inst]) ? p.high[inst] + (size_t)sidx * p.high_stream_stride[inst] + (size_t)f * p.high_frame_stride[inst] : nullptr;
    int16_t* out = live ? p.out[inst] + (size_t)sidx * p.out_stream_stride[inst] + (size_t)f * p.out_frame_stride[inst] : nullptr;
    for (int i0 = 0; i0 < p.band_len; i0 += 8) {
      uint4 lw = make_uint4(0u, 0u, 0u, 0u), hw = make_uint4(0u, 0u, 0u, 0u);
      if (lo) lw = *reinterpret_cast<const uint4*>(lo + i0);
      if (hi) hw = *reinterpret_cast<const uint4*>(hi + i0);
      const uint32_t l[4] = {lw.x, lw.y, lw.z, lw.w}, h[4] = {hw.x, hw.y, hw.z, hw.w};
      uint32_t ow[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int lv = (int)(int16_t)((k & 1) ? (l[k >> 1] >> 16) : (l[k >> 1] & 0xffffu));
        const int hv = (int)(int16_t)((k & 1) ? (h[k >> 1] >> 16) : (h[k >> 1] & 0xffffu));
        const int x = (chain == 0 ? lv + hv : lv - hv) * 1024;
        const int y = band_allpass3(x, st, c0, c1, c2);
        const int v = band_sat16((y + 512) >> 10);
        const int other = __shfl_xor_sync(kFullMask, v, 1);
        // out[2i] = chain 1 (filter2), out[2i+1] = chain 0 (filter1)
        const int even = chain == 0 ? other : v, odd = chain == 0 ? v : other;
        ow[k] = ((uint32_t)even & 0xffffu) | ((uint32_t)odd << 16);
      }
      // the pair of lanes holds the same 8 words: chain 0 stores the first half
      if (out) {
        if (chain == 0) *reinterpret_cast<uint4*>(out + 2 * i0) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        else *reinterpret_cast<uint4*>(out + 2 * i0 + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
      }
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

template <bool UP>
__global__ void __launch_bounds__(kResampleWarpsPerCta * 32)
resample_kernel(const ResampleLaunch p) {
  constexpr int SRC = UP ? 480 : 640;
  constexpr int DST = UP ? 640 : 480;
  constexpr int NOUT = DST / 32;         // outputs per lane: 20 / 15
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
  int16_t* ghist = reinterpret_cast<int16_t*>(gst + (UP ? kBandOffAnaHist : kBandOffSynHist));
  for (int i = lane; i < 64; i += 32) E[i] = (float)ghist[i];
  __syncwarp();

  for (int f = 0; f < p.frames; ++f) {
    const int16_t* in = p.in + (size_t)sidx * p.in_stream_stride + (size_t)f * p.in_frame_stride;
    int16_t* out = p.out + (size_t)sidx * p.out_stream_stride + (size_t)f * p.out_frame_stride;
    for (int w = lane; w < SRC / 2; w += 32) {
      const uint32_t v = reinterpret_cast<const uint32_t*>(in)[w];
      E[64 + 2 * w] = (float)(int16_t)(v & 0xffffu);
      E[64 + 2 * w + 1] = (float)(int16_t)(v >> 16);
    }
    __syncwarp();
    const int32_t* sch = UP ? nullptr : p.schedule + (size_t)f * 480 * 3;
#pragma unroll 1
    for (int t = 0; t < NOUT; ++t) {
      const int n = lane + 32 * t;
      int epos, off;
      float fac, fac1;
      if (UP) {
        // e_n = 31.5 + 0.75 n  ->  4 e_n = 126 + 3 n; remainders are multiples of 1/4,
        // so the table row is exact and the interpolation factor is 0
        const int q4 = 126 + 3 * n;
        epos = q4 >> 2;
        off = (q4 & 3) * 8;
        fac = 0.f;
        fac1 = 1.f;
      } else {
        const int pk = sch[3 * n];
        epos = pk >> 8;
        off = pk & 0xff;
        fac = __int_as_float(sch[3 * n + 1]);
        fac1 = __int_as_float(sch[3 * n + 2]);
      }
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
      const float res = (t0 + t2) + (t1 + t3);
      out[n] = (int16_t)band_round_s16(res);
    }
    __syncwarp();
    // history: last 64 input samples
    float h0 = E[SRC + lane], h1 = E[SRC + 32 + lane];
    __syncwarp();
    E[lane] = h0;
    E[32 + lane] = h1;
    __syncwarp();
  }
  for (int i = lane; i < 64; i += 32) ghist[i] = (int16_t)E[i];
}

// ---- launch helpers (host) -------------------------------------------------------
struct BandLaunch {
  int32_t* state;
  const int* slots;
  const float* kernel_up;     // resampler tables (device)
  const float* kernel_down;
  // 48 kHz merge: the 640 -> 480 resampler runs once per group of streams that share a
  // position schedule (normally one group = the whole batch)
  struct DownGroup { const int32_t* schedule; const int* stream_index; int count; };
  std::vector<DownGroup> down_groups;
  int16_t* full;              // full-band PCM [stream][frame][fs/100]
  long long full_stride;
  int16_t* bands;             // [stream][frame][nb][160]
  long long bands_stride;
  int16_t* scratch;           // 48 kHz: [2][stream][frame][640]
  int n_streams, frames;
};

inline size_t BandScratchElems(int nb, int n_streams, int frames) {
  return nb == 3 ? (size_t)2 * n_streams * frames * 640 : 0;
}

inline int LaunchBandSplit(int nb, const BandLaunch& b, cudaStream_t st, uint64_t* launches) {
  const int n = b.n_streams, F = b.frames;
  const long long bs = b.bands_stride, fstride = nb * 160;
  if (nb == 2) {
    QmfAnaLaunch q = {};
    q.state = b.state; q.slots = b.slots; q.n_streams = n; q.frames = F; q.instances = 1; q.len = 320;
    q.state_off[0] = kBandOffQmf0;
    q.in[0] = b.full; q.in_stream_stride[0] = b.full_stride; q.in_frame_stride[0] = 320;
    q.low[0] = b.bands; q.low_stream_stride[0] = bs; q.low_frame_stride[0] = fstride;
    q.high[0] = b.bands + 160; q.high_stream_stride[0] = bs; q.high_frame_stride[0] = fstride;
    qmf_analysis_kernel<<<(2 * n + 127) / 128, 128, 0, st>>>(q);
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  if (nb != 3) return -1;
  int16_t* s64 = b.scratch;                                  // resampled 64 kHz signal
  int16_t* s32 = b.scratch + (size_t)n * F * 640;            // [stream][frame][low 320 | high 320]
  const long long ss = (long long)F * 640;
  ResampleLaunch r = {};
  r.state = b.state; r.slots = b.slots; r.kernel = b.kernel_up; r.in = b.full; r.out = s64;
  r.in_stream_stride = b.full_stride; r.in_frame_stride = 480; r.out_stream_stride = ss; r.out_frame_stride = 640;
  r.n_streams = n; r.frames = F;
  const size_t smem = kResampleSmemBytes;
  resample_kernel<true><<<(n + kResampleWarpsPerCta - 1) / kResampleWarpsPerCta, kResampleWarpsPerCta * 32, smem, st>>>(r);
  ++*launches;
  QmfAnaLaunch q = {};
  q.state = b.state; q.slots = b.slots; q.n_streams = n; q.frames = F; q.instances = 1; q.len = 640;
  q.state_off[0] = kBandOffQmf0;
  q.in[0] = s64; q.in_stream_stride[0] = ss; q.in_frame_stride[0] = 640;
  q.low[0] = s32; q.low_stream_stride[0] = ss; q.low_frame_stride[0] = 640;
  q.high[0] = s32 + 320; q.high_stream_stride[0] = ss; q.high_frame_stride[0] = 640;
  qmf_analysis_kernel<<<(2 * n + 127) / 128, 128, 0, st>>>(q);
  ++*launches;
  QmfAnaLaunch q2 = {};
  q2.state = b.state; q2.slots = b.slots; q2.n_streams = n; q2.frames = F; q2.instances = 2; q2.len = 320;
  q2.state_off[0] = kBandOffQmf1; q2.state_off[1] = kBandOffQmf2;
  q2.in[0] = s32; q2.in[1] = s32 + 320;
  for (int i = 0; i < 2; ++i) { q2.in_stream_stride[i] = ss; q2.in_frame_stride[i] = 640; }
  q2.low[0] = b.bands; q2.high[0] = b.bands + 160;           // 0-8 kHz, 8-16 kHz
  q2.low[1] = nullptr; q2.high[1] = b.bands + 320;           // (24-32 kHz dropped), 16-24 kHz
  for (int i = 0; i < 2; ++i) {
    q2.low_stream_stride[i] = q2.high_stream_stride[i] = bs;
    q2.low_frame_stride[i] = q2.high_frame_stride[i] = fstride;
  }
  qmf_analysis_kernel<<<(4 * n + 127) / 128, 128, 0, st>>>(q2);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

inline int LaunchBandMerge(int nb, const BandLaunch& b, cudaStream_t st, uint64_t* launches) {
  const int n = b.n_streams, F = b.frames;
  const long long bs = b.bands_stride, fstride = nb * 160;
  if (nb == 2) {
    QmfSynLaunch q = {};
    q.state = b.state; q.slots = b.slots; q.n_streams = n; q.frames = F; q.instances = 1; q.band_len = 160;
    q.state_off[0] = kBandOffQmf0 + 12;
    q.low[0] = b.bands; q.high[0] = b.bands + 160;
    q.low_stream_stride[0] = q.high_stream_stride[0] = bs;
    q.low_frame_stride[0] = q.high_frame_stride[0] = fstride;
    q.out[0] = b.full; q.out_stream_stride[0] = b.full_stride; q.out_frame_stride[0] = 320;
    qmf_synthesis_kernel<<<(2 * n + 127) / 128, 128, 0, st>>>(q);
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  if (nb != 3) return -1;
  int16_t* s64 = b.scratch;
  int16_t* s32 = b.scratch + (size_t)n * F * 640;
  const long long ss = (long long)F * 640;
  QmfSynLaunch q2 = {};
  q2.state = b.state; q2.slots = b.slots; q2.n_streams = n; q2.frames = F; q2.instances = 2; q2.band_len = 160;
  q2.state_off[0] = kBandOffQmf1 + 12; q2.state_off[1] = kBandOffQmf2 + 12;
  q2.low[0] = b.bands; q2.high[0] = b.bands + 160;
  q2.low[1] = nullptr; q2.high[1] = b.bands + 320;           // zero low half (splitting_filter.cc:143-145,152)
  for (int i = 0; i < 2; ++i) {
    q2.low_stream_stride[i] = q2.high_stream_stride[i] = bs;
    q2.low_frame_stride[i] = q2.high_frame_stride[i] = fstride;
    q2.out_stream_stride[i] = ss; q2.out_frame_stride[i] = 640;
  }
  q2.out[0] = s32; q2.out[1] = s32 + 320;
  qmf_synthesis_kernel<<<(4 * n + 127) / 128, 128, 0, st>>>(q2);
  ++*launches;
  QmfSynLaunch q = {};
  q.state = b.state; q.slots = b.slots; q.n_streams = n; q.frames = F; q.instances = 1; q.band_len = 320;
  q.state_off[0] = kBandOffQmf0 + 12;
  q.low[0] = s32; q.high[0] = s32 + 320;
  q.low_stream_stride[0] = q.high_stream_stride[0] = ss;
  q.low_frame_stride[0] = q.high_frame_stride[0] = 640;
  q.out[0] = s64; q.out_stream_stride[0] = ss; q.out_frame_stride[0] = 640;
  qmf_synthesis_kernel<<<(2 * n + 127) / 128, 128, 0, st>>>(q);
  ++*launches;
  for (const BandLaunch::DownGroup& g : b.down_groups) {
    ResampleLaunch r = {};
    r.state = b.state; r.slots = b.slots; r.kernel = b.kernel_down; r.schedule = g.schedule;
    r.stream_index = g.stream_index; r.in = s64; r.out = b.full;
    r.in_stream_stride = ss; r.in_frame_stride = 640; r.out_stream_stride = b.full_stride; r.out_frame_stride = 480;
    r.n_streams = g.count; r.frames = F;
    resample_kernel<false><<<(g.count + kResampleWarpsPerCta - 1) / kResampleWarpsPerCta,
                             kResampleWarpsPerCta * 32, kResampleSmemBytes, st>>>(r);
    ++*launches;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace nsb200

#endif  // AUDIOSIGNALPROCESS_B200_BAND_KERNELS_CUH_
