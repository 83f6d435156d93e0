// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Thin driver over the UNMODIFIED reference sources (compiled where they lie
// under /root/reference by oracle/Makefile into oracle/_ref/libns_ref.so).
// Nothing in this file restates the algorithm: it only walks 10 ms frames
// through the reference's own public API, exactly as its three call sites do
// (WebRtc_AMP_Port/test_ns_module.cpp:83-109,
//  WebRtc_AMP_Port/libapm/src/apm_ns.cpp:96-132,
//  webrtc/modules/audio_processing/noise_suppression_impl.cc:66-95).
//
// Entry points (all extern "C", plain pointers):
//   ref_ns_run      float NS  (WebRtcNs_Analyze + WebRtcNs_Process), 8/16/32/48 kHz
//   ref_nsx_run     fixed NSx (WebRtcNsx_Process),                   8/16/32/48 kHz
//   ref_ns_run_mt / ref_nsx_run_mt   same over many streams with pthreads
//                   (the CPU baseline of bench.py)
//   ref_qmf_* / ref_resample_* / ref_spl_*  primitive hooks for the KATs.
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <vector>

#include "webrtc/common_audio/resampler/push_sinc_resampler.h"
#include "webrtc/common_audio/signal_processing/include/real_fft.h"
#include "webrtc/common_audio/signal_processing/include/signal_processing_library.h"
#include "libapm/include/apm_ns.h"
#include "webrtc/modules/audio_processing/audio_buffer.h"
#include "webrtc/modules/audio_processing/ns/include/noise_suppression.h"
#include "webrtc/modules/audio_processing/ns/include/noise_suppression_x.h"

using webrtc::AudioBuffer;

namespace {

inline int FrameLen(int fs) { return fs / 100; }

inline int16_t RoundToS16(float v) {  // audio_util.h:41-49 contract, via the reference's own helper
  return webrtc::FloatS16ToS16(v);
}

}  // namespace

extern "C" {

// Float NS over one stream. pcm_in: nframes*fs/100 int16. Outputs (any may be
// NULL):
//   out_f32    8/16 kHz only: the raw float frame WebRtcNs_Process wrote
//              (int16-scale floats, clamped, not rounded). Not written at
//              32/48 kHz, where AudioBuffer merges the bands in int16.
//   out_i16    the int16 frame a caller of AudioBuffer reads back.
//   prior_prob WebRtcNs_prior_speech_probability after each frame.
int ref_ns_run(int fs, int mode, int nframes, const int16_t* pcm_in,
               float* out_f32, int16_t* out_i16, float* prior_prob) {
  NsHandle* h = NULL;
  if (WebRtcNs_Create(&h) != 0) return -1;
  if (WebRtcNs_Init(h, (uint32_t)fs) != 0 || WebRtcNs_set_policy(h, mode) != 0) {
    WebRtcNs_Free(h);
    return -1;
  }
  const int n = FrameLen(fs);
  if (fs == 8000 || fs == 16000) {
    std::vector<float> in(n), out(n);
    for (int f = 0; f < nframes; ++f) {
      for (int i = 0; i < n; ++i) in[i] = (float)pcm_in[(size_t)f * n + i];
      const float* inb[1] = {in.data()};
      float* outb[1] = {out.data()};
      WebRtcNs_Analyze(h, in.data());
      WebRtcNs_Process(h, inb, 1, outb);
      for (int i = 0; i < n; ++i) {
        if (out_f32) out_f32[(size_t)f * n + i] = out[i];
        if (out_i16) out_i16[(size_t)f * n + i] = RoundToS16(out[i]);
      }
      if (prior_prob) prior_prob[f] = WebRtcNs_prior_speech_probability(h);
    }
  } else {
    AudioBuffer ab(n, 1, n, 1, n);
    for (int f = 0; f < nframes; ++f) {
      memcpy(ab.data(0), pcm_in + (size_t)f * n, sizeof(int16_t) * n);
      ab.SplitIntoFrequencyBands();
      WebRtcNs_Analyze(h, ab.split_bands_const_f(0)[webrtc::kBand0To8kHz]);
      WebRtcNs_Process(h, ab.split_bands_const_f(0), ab.num_bands(), ab.split_bands_f(0));
      ab.MergeFrequencyBands();
      if (out_i16) memcpy(out_i16 + (size_t)f * n, ab.data_const(0), sizeof(int16_t) * n);
      if (prior_prob) prior_prob[f] = WebRtcNs_prior_speech_probability(h);
    }
  }
  WebRtcNs_Free(h);
  return 0;
}

// Float NS with different signals for Analyze and Process (an echo canceller between them:
// audio_processing_impl.cc:625-631), straight through the reference API.  ana: [frame][fl] band-0
// frames for WebRtcNs_Analyze; in/out: [frame][band][fl] band frames for WebRtcNs_Process
// (int16-scale floats).  The first fused_frames frames feed band 0 of `in` to both.
int ref_ns_split_run(int fs, int mode, int nb, int nframes, int fused_frames, const float* ana,
                     const float* in, float* out) {
  NsHandle* h = NULL;
  if (WebRtcNs_Create(&h) != 0) return -1;
  if (WebRtcNs_Init(h, (uint32_t)fs) != 0 || WebRtcNs_set_policy(h, mode) != 0) {
    WebRtcNs_Free(h);
    return -1;
  }
  const int fl = fs == 8000 ? 80 : 160;
  for (int f = 0; f < nframes; ++f) {
    const float* bands_in[3];
    float* bands_out[3];
    for (int b = 0; b < nb; ++b) {
      bands_in[b] = in + ((size_t)f * nb + b) * fl;
      bands_out[b] = out + ((size_t)f * nb + b) * fl;
    }
    WebRtcNs_Analyze(h, f < fused_frames ? bands_in[0] : ana + (size_t)f * fl);
    WebRtcNs_Process(h, bands_in, nb, bands_out);
  }
  WebRtcNs_Free(h);
  return 0;
}

// Fixed NSx over one stream (explicit band split at 32/48 kHz, as
// noise_suppression_impl.cc:90-93 would do; the author's -DNS_FIXED driver
// never splits -- SURVEY.md appendix B.9).
int ref_nsx_run(int fs, int mode, int nframes, const int16_t* pcm_in, int16_t* out_i16) {
  NsxHandle* h = NULL;
  if (WebRtcNsx_Create(&h) != 0) return -1;
  if (WebRtcNsx_Init(h, (uint32_t)fs) != 0 || WebRtcNsx_set_policy(h, mode) != 0) {
    WebRtcNsx_Free(h);
    return -1;
  }
  const int n = FrameLen(fs);
  if (fs == 8000 || fs == 16000) {
    std::vector<int16_t> in(n), out(n);
    for (int f = 0; f < nframes; ++f) {
      memcpy(in.data(), pcm_in + (size_t)f * n, sizeof(int16_t) * n);
      const short* inb[1] = {in.data()};
      short* outb[1] = {out.data()};
      WebRtcNsx_Process(h, inb, 1, outb);
      memcpy(out_i16 + (size_t)f * n, out.data(), sizeof(int16_t) * n);
    }
  } else {
    AudioBuffer ab(n, 1, n, 1, n);
    for (int f = 0; f < nframes; ++f) {
      memcpy(ab.data(0), pcm_in + (size_t)f * n, sizeof(int16_t) * n);
      ab.SplitIntoFrequencyBands();
      WebRtcNsx_Process(h, ab.split_bands_const(0), ab.num_bands(), ab.split_bands(0));
      ab.MergeFrequencyBands();
      memcpy(out_i16 + (size_t)f * n, ab.data_const(0), sizeof(int16_t) * n);
    }
  }
  WebRtcNsx_Free(h);
  return 0;
}

// ---- many streams, pthreads: one stream per core at a time ------------------
struct MtJob {
  int fixed, fs, mode, nframes, nstreams, tid, nthreads;
  const int16_t* in;
  int16_t* out;
};

static void* MtWorker(void* p) {
  MtJob* j = (MtJob*)p;
  const size_t per = (size_t)j->nframes * FrameLen(j->fs);
  for (int s = j->tid; s < j->nstreams; s += j->nthreads) {
    if (j->fixed)
      ref_nsx_run(j->fs, j->mode, j->nframes, j->in + s * per, j->out + s * per);
    else
      ref_ns_run(j->fs, j->mode, j->nframes, j->in + s * per, NULL, j->out + s * per, NULL);
  }
  return NULL;
}

// pcm_in/out: [nstreams][nframes*fs/100] int16. Returns wall seconds spent in
// the processing loop only (CLOCK_MONOTONIC), or <0 on error.
double ref_run_mt(int fixed, int fs, int mode, int nstreams, int nframes, int nthreads,
                  const int16_t* pcm_in, int16_t* pcm_out) {
  if (nthreads < 1) nthreads = 1;
  std::vector<pthread_t> th(nthreads);
  std::vector<MtJob> jobs(nthreads);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; ++t) {
    MtJob j = {fixed, fs, mode, nframes, nstreams, t, nthreads, pcm_in, pcm_out};
    jobs[t] = j;
    if (pthread_create(&th[t], NULL, MtWorker, &jobs[t]) != 0) return -1.0;
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

// ---- primitive hooks for known-answer tests ---------------------------------
// state: 4 x 6 int32 = analysis_state1, analysis_state2, synthesis_state1, synthesis_state2
void ref_qmf_analysis(const int16_t* in, int in_len, int16_t* low, int16_t* high, int32_t* state) {
  WebRtcSpl_AnalysisQMF(in, in_len, low, high, state, state + 6);
}
void ref_qmf_synthesis(const int16_t* low, const int16_t* high, int band_len, int16_t* out,
                       int32_t* state) {
  WebRtcSpl_SynthesisQMF(low, high, band_len, out, state + 12, state + 18);
}

void* ref_resampler_create(int src_frames, int dst_frames) {
  return new webrtc::PushSincResampler(src_frames, dst_frames);
}
void ref_resampler_free(void* r) { delete (webrtc::PushSincResampler*)r; }
int ref_resampler_run(void* r, const int16_t* src, int src_frames, int16_t* dst, int dst_cap) {
  return ((webrtc::PushSincResampler*)r)->Resample(src, src_frames, dst, dst_cap);
}

// Band split / merge exactly as AudioBuffer does it (persistent object so the
// filter states carry across frames).
void* ref_split_create(int fs) {
  const int n = FrameLen(fs);
  return new AudioBuffer(n, 1, n, 1, n);
}
void ref_split_free(void* p) { delete (AudioBuffer*)p; }
// bands: [3][160] int16 (unused bands untouched). Returns num_bands.
int ref_split_analysis(void* p, const int16_t* in, int n, int16_t* bands) {
  AudioBuffer* ab = (AudioBuffer*)p;
  memcpy(ab->data(0), in, sizeof(int16_t) * n);
  ab->SplitIntoFrequencyBands();
  for (int b = 0; b < ab->num_bands(); ++b)
    memcpy(bands + 160 * b, ab->split_bands_const(0)[b], sizeof(int16_t) * 160);
  return ab->num_bands();
}
void ref_split_synthesis(void* p, const int16_t* bands, int n, int16_t* out) {
  AudioBuffer* ab = (AudioBuffer*)p;
  for (int b = 0; b < ab->num_bands(); ++b)
    memcpy(ab->split_bands(0)[b], bands + 160 * b, sizeof(int16_t) * 160);
  ab->MergeFrequencyBands();
  memcpy(out, ab->data_const(0), sizeof(int16_t) * n);
}

// The author's wrapper class (libapm/src/apm_ns.cpp), one AudioBuffer frame per call.
void* ref_apm_ns_create(unsigned fs, int mode, int frame, int channels) {
  APM_NS* a = new APM_NS();
  if (!a->initNsModule(fs, mode, frame, channels)) {
    delete a;
    return NULL;
  }
  return a;
}
void ref_apm_ns_free(void* a) { delete (APM_NS*)a; }
void ref_apm_ns_process_i16(void* a, int16_t* data, int frame, int channels) {
  ((APM_NS*)a)->processCaptureStream(data, frame, channels);
}
void ref_apm_ns_process_f32(void* a, float* data, int frame, int channels) {
  ((APM_NS*)a)->processCaptureStream(data, frame, channels);
}

// SPL scalar helpers (signal_processing_unittest.cc KAT values).
int32_t ref_spl_sqrt_floor(int32_t v) { return WebRtcSpl_SqrtFloor(v); }
int32_t ref_spl_energy(int16_t* v, int len, int* scale) { return WebRtcSpl_Energy(v, len, scale); }
int16_t ref_spl_norm_w32(int32_t a) { return WebRtcSpl_NormW32(a); }
int16_t ref_spl_norm_u32(uint32_t a) { return WebRtcSpl_NormU32(a); }
int16_t ref_spl_norm_w16(int16_t a) { return WebRtcSpl_NormW16(a); }
int16_t ref_spl_get_size_in_bits(uint32_t a) { return WebRtcSpl_GetSizeInBits(a); }
int32_t ref_spl_div_w32w16(int32_t n, int16_t d) { return WebRtcSpl_DivW32W16(n, d); }
uint32_t ref_spl_div_u32u16(uint32_t n, uint16_t d) { return WebRtcSpl_DivU32U16(n, d); }
int16_t ref_spl_max_abs_w16(const int16_t* v, int len) { return WebRtcSpl_MaxAbsValueW16(v, len); }

// order-`order` real FFT pair on int16 (real_fft.c:47,74). fwd: in[1<<order]
// -> out[(1<<order)+2]; inv: in[(1<<order)+2] -> out[1<<order], returns scale.
int ref_spl_real_fft(int order, int inverse, const int16_t* in, int16_t* out) {
  WebRtcSpl_Init();
  struct RealFFT* f = WebRtcSpl_CreateRealFFT(order);
  if (!f) return -1;
  int r = inverse ? WebRtcSpl_RealInverseFFT(f, in, out) : WebRtcSpl_RealForwardFFT(f, in, out);
  WebRtcSpl_FreeRealFFT(f);
  return r;
}

}  // extern "C"
